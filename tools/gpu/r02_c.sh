#!/bin/bash
# r02 GPU call C: coalesced-epilogue pair kernels + |x| prepass: unit tests, per-kernel ncu times, micro-bench, bench, full suite
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity_full.py -x -q -m gpu -k "persistent_pair or zplus_zero or teacher" > gpurun_out/c_unit.log 2>&1
echo "unit exit $?" >> gpurun_out/c_unit.log
timeout 600 ncu --metrics gpu__time_duration.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.sum.per_second --clock-control none -k regex:"te_tc_pair|abs_tf32" --csv --log-file gpurun_out/c_pair_times.csv python tools/profile_pair.py fc2 fc1 qkv proj > gpurun_out/c_ncu.log 2>&1
timeout 300 python tools/bench_kernels.py > gpurun_out/c_kern.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_default.json 2> gpurun_out/c_bench_default.err
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/c_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest_gpu.log
