#!/bin/bash
# r02 GPU call K (= H + J): full suite at the new defaults, mixed-kind GEMM unit test, A/B benches, launch list
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s -k "mixed_kind or 3xtf32 or attention_contractions or persistent_pair" > gpurun_out/k_unit.log 2>&1
echo "unit exit $?" >> gpurun_out/k_unit.log
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/k_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/k_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/k_bench_default.json 2> gpurun_out/k_bench_default.err
TE_B200_ATTN_PERSISTENT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/k_bench_noattnp.json 2> gpurun_out/k_bench_noattnp.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline --flags 307 > gpurun_out/k_bench_307.json 2> gpurun_out/k_bench_307.err
TE_B200_LINEAR_MIXED=1 timeout 300 python tools/bench_kernels.py > gpurun_out/k_kern_mixed.log 2>&1
TE_B200_LINEAR_MIXED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/k_bench_mixed.json 2> gpurun_out/k_bench_mixed.err
TE_B200_LINEAR_MIXED=1 timeout 1200 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_vit.py -q -m gpu -s -k "conditioned or golden or tiny" > gpurun_out/k_parity_mixed.log 2>&1
echo "parity exit $?" >> gpurun_out/k_parity_mixed.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/k_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/k_launches_bench.log 2>&1
