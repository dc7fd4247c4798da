#!/bin/bash
# dense rollout (aggregation + tcgen05 N^3 chain) at several batch sizes: fraction of the HBM copy peak
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/x6_dense_rollout.log 2>&1 <<'PY'
import json, torch, bench
from transformer_explainability_b200 import _lib
pk = bench.peaks()
w = bench.WORKLOADS["vit_base"]
for B in (32, 64, 128, 256):
    r = bench.roofline_rollout(w, _lib.FLAG_BENCH_DEFAULT, pk, B=B, dense=True)
    print(json.dumps({k: r[k] for k in ("kernel", "achieved", "frac", "ms")}), flush=True)
    torch.cuda.empty_cache()
w = bench.WORKLOADS["bert_base"]
for B in (16, 32, 64):
    r = bench.roofline_rollout(w, _lib.FLAG_BENCH_DEFAULT, pk, B=B, dense=True)
    print(json.dumps({k: r[k] for k in ("kernel", "achieved", "frac", "ms")}), flush=True)
    torch.cuda.empty_cache()
PY
cat gpurun_out/x6_dense_rollout.log
