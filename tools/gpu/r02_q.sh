#!/bin/bash
# fp16 R kernel (TE_FLAG_ZPLUS_R_F16): unit tests, microbench, bench A/B 7475 vs 15667
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "f16_split or fp16_second or persistent_pair" -s > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"
grep "rows .* in\|fp16 R" gpurun_out/q_pytest.log; tail -3 gpurun_out/q_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity_full.py -x -q -s -k "vit_base or bert" > gpurun_out/q_parity.log 2>&1; echo "parity rc=$?"
grep "flags" gpurun_out/q_parity.log | grep "7475\|15667" | cut -c1-260; tail -3 gpurun_out/q_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 --flags 15667 --no-cpu-baseline --no-roofline > gpurun_out/q_bench_15667.json 2> gpurun_out/q_bench_15667.err; tail -c 300 gpurun_out/q_bench_15667.json
timeout 600 python bench.py --steps 5 --warmup 3 --flags 7475 --no-cpu-baseline --no-roofline > gpurun_out/q_bench_7475.json 2> gpurun_out/q_bench_7475.err; tail -c 300 gpurun_out/q_bench_7475.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 1 --warmup 1 --flags 15667 --no-cpu-baseline --no-roofline > gpurun_out/q_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/q_launches.csv > gpurun_out/q_launches.txt; head -24 gpurun_out/q_launches.txt
