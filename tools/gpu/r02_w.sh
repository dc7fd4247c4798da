#!/bin/bash
# fast safe_divide in the S kernel, fast GELU' in the TF32 backward epilogue, back-off waits in the attention kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_rules.py tests/test_gpu_vit.py tests/test_gpu_bert.py -x -q > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/w_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity_full.py -x -q -s -k "vit_base or conditioned_bert" > gpurun_out/w_parity.log 2>&1; echo "parity rc=$?"
grep "flags" gpurun_out/w_parity.log | grep "flags 307:\|7475" | cut -c1-250; tail -2 gpurun_out/w_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/w_bench_default.json 2> gpurun_out/w_bench_default.err; tail -c 250 gpurun_out/w_bench_default.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/w_bench_default2.json 2> gpurun_out/w_bench_default2.err; tail -c 250 gpurun_out/w_bench_default2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/w_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/w_launches.csv > gpurun_out/w_launches.txt; head -30 gpurun_out/w_launches.txt
