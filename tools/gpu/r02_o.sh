#!/bin/bash
# fp16-split forward Linear: unit tests, kernel microbench, ncu launch list, full bench A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "f16_split or 3xtf32" -s > gpurun_out/o_pytest_f16.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/o_pytest_f16.log
timeout 300 python tools/bench_kernels.py > gpurun_out/o_kernels.log 2>&1; grep "linear fwd" gpurun_out/o_kernels.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/o_fwd16_launches.csv python tools/profile_fwd16.py > gpurun_out/o_profile.log 2>&1
tail -5 gpurun_out/o_profile.log
timeout 600 python bench.py --steps 5 --warmup 3 --flags 7475 --no-cpu-baseline > gpurun_out/o_bench_7475.json 2> gpurun_out/o_bench_7475.err; tail -c 600 gpurun_out/o_bench_7475.json
timeout 600 python bench.py --steps 5 --warmup 3 --flags 3379 --no-cpu-baseline > gpurun_out/o_bench_3379.json 2> gpurun_out/o_bench_3379.err; tail -c 300 gpurun_out/o_bench_3379.json
