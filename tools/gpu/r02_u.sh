#!/bin/bash
# cheaper softmax / safe_divide epilogues of the N x N attention kernels: attention + engine tests, parity (ViT-B, BERT), bench, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_rules.py tests/test_gpu_vit.py -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity_full.py -x -q -s -k "vit_base or conditioned_bert" > gpurun_out/u_parity.log 2>&1; echo "parity rc=$?"
grep "flags" gpurun_out/u_parity.log | grep "flags 0:\|flags 51:\|7475" | cut -c1-250; tail -2 gpurun_out/u_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/u_bench_default.json 2> gpurun_out/u_bench_default.err; tail -c 250 gpurun_out/u_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/u_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/u_launches.csv > gpurun_out/u_launches.txt; grep "attn_nn" gpurun_out/u_launches.txt
