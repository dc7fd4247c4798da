#!/bin/bash
# r02 GPU call G: NK single-pass backward, one-launch dense aggregation: full suite, bench, launch list
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/g_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_pytest_gpu.log
grep -E "conditioned flags" gpurun_out/g_pytest_gpu.log > /dev/null
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/g_launches_bench.log 2>&1
