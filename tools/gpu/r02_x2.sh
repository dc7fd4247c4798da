#!/bin/bash
# r02 final validation of the final tree: full GPU suite, smoke(), bench lines of every workload (default flags 7475), reference arm, launch list
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/x2_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/x2_pytest_gpu.log; tail -3 gpurun_out/x2_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/x2_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/x2_bench_default.json 2> gpurun_out/x2_bench_default.err; tail -c 300 gpurun_out/x2_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/x2_bench_reference.json 2> gpurun_out/x2_bench_reference.err; tail -c 200 gpurun_out/x2_bench_reference.json
timeout 900 python bench.py --workload bert_base --steps 3 --warmup 3 --cpu-samples 3 > gpurun_out/x2_bench_bert.json 2> gpurun_out/x2_bench_bert.err; tail -c 200 gpurun_out/x2_bench_bert.json
timeout 900 python bench.py --workload vit_large --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/x2_bench_vitl.json 2> gpurun_out/x2_bench_vitl.err; tail -c 200 gpurun_out/x2_bench_vitl.json
timeout 900 python bench.py --workload deit_base_distilled --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/x2_bench_deit.json 2> gpurun_out/x2_bench_deit.err; tail -c 200 gpurun_out/x2_bench_deit.json
timeout 600 python bench.py --scaling strong --batch 32 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/x2_bench_b32.json 2> gpurun_out/x2_bench_b32.err; tail -c 200 gpurun_out/x2_bench_b32.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/x2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/x2_launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/x2_launches.csv > gpurun_out/x2_launches.txt; head -8 gpurun_out/x2_launches.txt
