#!/bin/bash
# r02 GPU call B: full GPU suite (no -x), ncu of the persistent pair kernels, launch list of the default bench step
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/b_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:te_tc_pair -o gpurun_out/prof_pair python tools/profile_pair.py fc2 fc1 > gpurun_out/b_ncu_pair.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/b_launches_bench.log 2>&1
timeout 900 python bench.py --workload bert_base --steps 3 --warmup 2 --no-cpu-baseline --no-roofline > gpurun_out/b_bench_bert.json 2> gpurun_out/b_bench_bert.err
