#!/bin/bash
# r02 GPU call L: persistent pair mixed GEMM, bf16 S1 denominator, old pair 3x: unit tests + A/B benches + parity
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s -k "mixed_kind or bf16_single_pass" > gpurun_out/l_unit.log 2>&1
echo "unit exit $?" >> gpurun_out/l_unit.log
TE_B200_LINEAR_MIXED=2 timeout 300 python tools/bench_kernels.py > gpurun_out/l_kern_mixedpair.log 2>&1
TE_B200_LINEAR_2CTA=1 timeout 300 python tools/bench_kernels.py > gpurun_out/l_kern_3xpair.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/l_bench_default.json 2> gpurun_out/l_bench_default.err
TE_B200_LINEAR_MIXED=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/l_bench_mixedpair.json 2> gpurun_out/l_bench_mixedpair.err
TE_B200_LINEAR_MIXED=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline --flags 3379 > gpurun_out/l_bench_mixedpair_bf16s1.json 2> gpurun_out/l_bench_mixedpair_bf16s1.err
TE_B200_LINEAR_MIXED=2 timeout 1200 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_vit.py tests/test_gpu_bert.py -q -m gpu -s -k "conditioned or golden or tiny" > gpurun_out/l_parity_mixedpair.log 2>&1
echo "parity exit $?" >> gpurun_out/l_parity_mixedpair.log
