#!/bin/bash
# r02 GPU call N (--gpus G): weak + strong scaling line at G GPUs of one box (the strong numbers ride in the weak line's "strong" key)
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
G=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/n_smi_$G.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/n_bench_${G}gpu.json 2> gpurun_out/n_bench_${G}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $G --steps 5 --warmup 3 --impl reference > gpurun_out/n_bench_${G}gpu_reference.json 2> gpurun_out/n_bench_${G}gpu_reference.err
