#!/bin/bash
# fp16 single-pass kernels (R kernel fed by the S kernel's fp16 epilogue; backward Linears): tests, bench A/B, launch list
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "fp16_second or fp16_single_pass or f16_split or persistent_pair or bf16_single" -s > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"
grep "fp16 R\|fp16 backward" gpurun_out/r_pytest.log; tail -3 gpurun_out/r_pytest.log
for f in 15667 23859 32051 7475; do
timeout 600 python bench.py --steps 5 --warmup 3 --flags $f --no-cpu-baseline --no-roofline > gpurun_out/r_bench_$f.json 2> gpurun_out/r_bench_$f.err; tail -c 200 gpurun_out/r_bench_$f.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_launches.csv python bench.py --steps 1 --warmup 1 --flags 32051 --no-cpu-baseline --no-roofline > gpurun_out/r_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r_launches.csv > gpurun_out/r_launches.txt; head -24 gpurun_out/r_launches.txt
