#!/bin/bash
# r02 GPU call F: default bench with every roofline + CPU arm on the reference mirror; reference arm; BERT + ViT-L + DeiT lines
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
timeout 900 python bench.py --workload bert_base --steps 3 --warmup 3 --cpu-samples 3 > gpurun_out/f_bench_bert.json 2> gpurun_out/f_bench_bert.err
timeout 900 python bench.py --workload vit_large --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_vitl.json 2> gpurun_out/f_bench_vitl.err
timeout 900 python bench.py --workload deit_base_distilled --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_deit.json 2> gpurun_out/f_bench_deit.err
