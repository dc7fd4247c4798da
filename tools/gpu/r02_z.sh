#!/bin/bash
# pipelined x loads in the R kernel epilogue, leaner S / GELU' epilogues: rule + engine tests, parity, bench x2, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_rules.py tests/test_gpu_vit.py tests/test_gpu_bert.py -x -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/z_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity_full.py -x -q -s -k "vit_base or conditioned_bert or teacher or zero" > gpurun_out/z_parity.log 2>&1; echo "parity rc=$?"
grep "flags" gpurun_out/z_parity.log | grep "7475" | cut -c1-250; tail -2 gpurun_out/z_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/z_bench_default.json 2> gpurun_out/z_bench_default.err; tail -c 250 gpurun_out/z_bench_default.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/z_bench_default2.json 2> gpurun_out/z_bench_default2.err; tail -c 250 gpurun_out/z_bench_default2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/z_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/z_launches.csv > gpurun_out/z_launches.txt; head -10 gpurun_out/z_launches.txt
