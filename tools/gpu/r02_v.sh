#!/bin/bash
# fused GELU split emission (232-register build): A/B against the pre-pass, parity of the forward, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_full.py -x -q -s -k "vit_base or conditioned_bert" > gpurun_out/v_parity.log 2>&1; echo "parity rc=$?"
grep "flags" gpurun_out/v_parity.log | grep "flags 51:\|4147\|7475" | cut -c1-250; tail -2 gpurun_out/v_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/v_bench_fused.json 2> gpurun_out/v_bench_fused.err; tail -c 250 gpurun_out/v_bench_fused.json
TE_B200_GELU_SPLIT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/v_bench_prepass.json 2> gpurun_out/v_bench_prepass.err; tail -c 250 gpurun_out/v_bench_prepass.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/v_bench_fused2.json 2> gpurun_out/v_bench_fused2.err; tail -c 250 gpurun_out/v_bench_fused2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/v_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/v_launches.csv > gpurun_out/v_launches.txt; head -14 gpurun_out/v_launches.txt; grep "blocksplit\|layernorm" gpurun_out/v_launches.txt
