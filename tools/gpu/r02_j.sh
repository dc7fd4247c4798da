#!/bin/bash
# r02 GPU call J: mixed-kind forward GEMM + persistent N x N attention kernel: unit tests, A/B benches
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -s -k "mixed_kind or 3xtf32 or attention_contractions" > gpurun_out/j_unit.log 2>&1
echo "unit exit $?" >> gpurun_out/j_unit.log
TE_B200_LINEAR_MIXED=1 timeout 300 python tools/bench_kernels.py > gpurun_out/j_kern_mixed.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/j_bench_default.json 2> gpurun_out/j_bench_default.err
TE_B200_ATTN_PERSISTENT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/j_bench_noattnp.json 2> gpurun_out/j_bench_noattnp.err
TE_B200_LINEAR_MIXED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/j_bench_mixed.json 2> gpurun_out/j_bench_mixed.err
TE_B200_LINEAR_MIXED=1 timeout 1200 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_vit.py -q -m gpu -s -k "conditioned or golden or tiny" > gpurun_out/j_parity_mixed.log 2>&1
echo "parity exit $?" >> gpurun_out/j_parity_mixed.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/j_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/j_launches_bench.log 2>&1
