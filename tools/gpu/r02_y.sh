#!/bin/bash
# per-kernel issue utilisation / tensor activity / DRAM throughput over one bench step (which kernels are instruction-bound?)
set -x
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -c 3500 --csv --log-file gpurun_out/y_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/y_ncu.log 2>&1
python tools/summarize_metrics.py gpurun_out/y_metrics.csv > gpurun_out/y_metrics.txt; head -40 gpurun_out/y_metrics.txt | cut -c1-200
