#!/bin/bash
# last check of the committed tree: full GPU suite (incl. the bit-exact block-split format test), smoke, one default bench line
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/x3_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/x3_pytest_gpu.log; tail -3 gpurun_out/x3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x3_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/x3_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/x3_bench_default.json 2> gpurun_out/x3_bench_default.err; tail -c 300 gpurun_out/x3_bench_default.json
