#!/bin/bash
# r02 GPU call D: S1 epilogue fix, coalesced epilogues in gemm3x / attn_nn: tests, micro-bench, bench, launch list
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_rules.py tests/test_gpu_parity_full.py -x -q -m gpu -k "not bert" > gpurun_out/d_unit.log 2>&1
echo "unit exit $?" >> gpurun_out/d_unit.log
timeout 300 python tools/bench_kernels.py > gpurun_out/d_kern.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_default.json 2> gpurun_out/d_bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/d_launches_bench.log 2>&1
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/d_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest_gpu.log
timeout 600 python bench.py --workload bert_base --steps 3 --warmup 2 --no-cpu-baseline --no-roofline > gpurun_out/d_bench_bert.json 2> gpurun_out/d_bench_bert.err
timeout 600 python bench.py --workload vit_large --steps 3 --warmup 2 --no-cpu-baseline --no-roofline > gpurun_out/d_bench_vitl.json 2> gpurun_out/d_bench_vitl.err
