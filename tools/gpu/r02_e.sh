#!/bin/bash
# r02 GPU call E: layers_lrp variant, hdf5 driver, graph replay, truncation compensation: full suite + bench
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/e_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/e_bench_default.json 2> gpurun_out/e_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/e_bench_reference.json 2> gpurun_out/e_bench_reference.err
timeout 600 python bench.py --scaling strong --batch 32 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/e_bench_b32.json 2> gpurun_out/e_bench_b32.err
