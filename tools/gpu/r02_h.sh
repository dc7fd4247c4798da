#!/bin/bash
# r02 GPU call H: single-pass NN / relevance NK kernels (flag 1024): full suite, bench at 307 and default, launch list
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/h_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/h_bench_default.json 2> gpurun_out/h_bench_default.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline --flags 307 > gpurun_out/h_bench_307.json 2> gpurun_out/h_bench_307.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/h_launches_bench.log 2>&1
