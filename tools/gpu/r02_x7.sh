#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/x7_bench_default.json 2> gpurun_out/x7_bench_default.err; tail -c 600 gpurun_out/x7_bench_default.json
