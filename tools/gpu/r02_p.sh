#!/bin/bash
# fp16-split forward Linear with fused split emission: unit tests, parity tables, launch list, bench A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "f16_split" -s > gpurun_out/p_pytest_f16.log 2>&1; echo "pytest rc=$?"
grep "rows .* in" gpurun_out/p_pytest_f16.log; tail -3 gpurun_out/p_pytest_f16.log
timeout 1500 python -m pytest tests/test_gpu_parity_full.py -x -q -s > gpurun_out/p_pytest_parity.log 2>&1; echo "parity rc=$?"
grep "flags" gpurun_out/p_pytest_parity.log | cut -c1-260; tail -3 gpurun_out/p_pytest_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_default.json 2> gpurun_out/p_bench_default.err; tail -c 400 gpurun_out/p_bench_default.json
timeout 600 python bench.py --steps 5 --warmup 3 --flags 3379 --no-cpu-baseline --no-roofline > gpurun_out/p_bench_3379.json 2> gpurun_out/p_bench_3379.err; tail -c 300 gpurun_out/p_bench_3379.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/p_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/p_launches.csv > gpurun_out/p_launches.txt; head -30 gpurun_out/p_launches.txt
