#!/bin/bash
# r02 GPU call A: new persistent pair kernels (unit tests under a short timeout first), full GPU suite, kernel micro-bench
# old vs new, bench default vs r01 selection.
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu -k "persistent_pair" > gpurun_out/a_pair_unit.log 2>&1
echo "pair unit exit $?" >> gpurun_out/a_pair_unit.log
timeout 300 python tools/bench_kernels.py > gpurun_out/a_kern_new.log 2>&1
TE_B200_ZPLUS_PERSISTENT=0 timeout 300 python tools/bench_kernels.py > gpurun_out/a_kern_old.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --flags 51 > gpurun_out/a_bench_51_persistent.json 2> gpurun_out/a_bench_51.err
TE_B200_ZPLUS_PERSISTENT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline --flags 51 > gpurun_out/a_bench_51_r01.json 2> gpurun_out/a_bench_51_r01.err
timeout 2400 python -m pytest tests -q -m gpu -x -s > gpurun_out/a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest_gpu.log
timeout 900 python tools/diag_flags.py 32 > gpurun_out/a_flag_bisect.log 2>&1
