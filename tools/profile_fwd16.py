"""One call of every forward-Linear kernel variant at the bench shapes — run under `ncu --metrics gpu__time_duration.sum` to get
the per-kernel times (row-split pre-pass and GEMM separately), or under `ncu --set full -k regex:fwd16` for the full capture."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformer_explainability_b200 import ops      # noqa: E402

rows = int(os.environ.get("ROWS", 256 * 197))
g = torch.Generator(device="cuda").manual_seed(1)
for name, inf, outf in [("fc2", 3072, 768), ("fc1", 768, 3072), ("qkv", 768, 2304), ("proj", 768, 768)]:
    x = torch.randn(rows, inf, device="cuda", generator=g)
    w = torch.randn(outf, inf, device="cuda", generator=g) * 0.02
    b = torch.randn(outf, device="cuda", generator=g) * 0.02
    for rep in range(2):
        y3 = ops.linear_forward(x, w, b, tensor_cores=True)
        yh = ops.linear_forward(x, w, b, tensor_cores=True, f16_split=True)
    torch.cuda.synchronize()
    print(name, "max |fp16split - 3xTF32| / max|y| = %.2e" % ((yh - y3).abs().max() / y3.abs().max()).item(), flush=True)
