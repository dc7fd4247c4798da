#!/bin/bash
# ncu --set full captures: N x N attention kernels (softmax / safe_divide epilogues) and the fp16-split forward GEMM
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:te_tc_attn_nn_kernel<4' -s 14 -c 2 -f -o gpurun_out/s_attn_softmax python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/s_ncu1.log 2>&1; tail -2 gpurun_out/s_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:te_tc_attn_nn_kernel<2' -s 14 -c 2 -f -o gpurun_out/s_attn_sd python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/s_ncu2.log 2>&1; tail -2 gpurun_out/s_ncu2.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:te_tc_fwd16_kernel' -s 50 -c 4 -f -o gpurun_out/s_fwd16 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/s_ncu3.log 2>&1; tail -2 gpurun_out/s_ncu3.log
ls -la gpurun_out/s_*.ncu-rep
