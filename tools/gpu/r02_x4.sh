#!/bin/bash
# ncu --set full of the final fp16-split forward kernel (232-register build): qkv / proj / fc1 / fc2 launches of one full-batch pass
set -x
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:te_tc_fwd16_kernel -s 60 -c 4 -f -o gpurun_out/x4_fwd16 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/x4_ncu.log 2>&1; tail -2 gpurun_out/x4_ncu.log
