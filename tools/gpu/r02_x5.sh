#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2 (fp16-split GEMM modes, split pre-passes / emissions, S-kernel fp16 epilogue) and smoke()
set -x
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tc.py -x -q -k "block_split or (f16_split and 394) or (fp16_second and 394) or (fp16_single and 394) or (persistent_pair and 394)" > gpurun_out/x5_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -6 gpurun_out/x5_memcheck_tests.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x5_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -5 gpurun_out/x5_memcheck_smoke.log
