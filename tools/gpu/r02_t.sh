#!/bin/bash
# register reallocation (setmaxnreg) in the fp16-split forward kernel: tests, bench, launch list; ncu of the N x N attention kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "f16_split or fp16_second or fp16_single_pass" -s > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"
grep "rows .* in" gpurun_out/t_pytest.log | head -8; tail -2 gpurun_out/t_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench_default.json 2> gpurun_out/t_bench_default.err; tail -c 300 gpurun_out/t_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/t_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/t_launches.csv > gpurun_out/t_launches.txt; head -16 gpurun_out/t_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:te_tc_attn_nn_kernel<\(int\)4' -s 14 -c 2 -f -o gpurun_out/t_attn_softmax python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/t_ncu1.log 2>&1; tail -2 gpurun_out/t_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:te_tc_attn_nn_kernel<\(int\)2' -s 14 -c 2 -f -o gpurun_out/t_attn_sd python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/t_ncu2.log 2>&1; tail -2 gpurun_out/t_ncu2.log
ls -la gpurun_out/t_*.ncu-rep
