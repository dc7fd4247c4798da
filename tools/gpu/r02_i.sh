#!/bin/bash
# r02 GPU call I: mixed-kind forward GEMM: unit test, micro-bench with and without, step bench with and without
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s -k "mixed_kind or 3xtf32" > gpurun_out/i_unit.log 2>&1
echo "unit exit $?" >> gpurun_out/i_unit.log
TE_B200_LINEAR_MIXED=1 timeout 300 python tools/bench_kernels.py > gpurun_out/i_kern_mixed.log 2>&1
TE_B200_LINEAR_MIXED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/i_bench_mixed.json 2> gpurun_out/i_bench_mixed.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/i_bench_3x.json 2> gpurun_out/i_bench_3x.err
TE_B200_LINEAR_MIXED=1 timeout 1200 python -m pytest tests/test_gpu_parity_full.py -q -m gpu -s -k "conditioned" > gpurun_out/i_parity_mixed.log 2>&1
echo "parity exit $?" >> gpurun_out/i_parity_mixed.log
