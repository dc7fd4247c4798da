"""GPU diagnostic: per-block error of the engine vs the fp64 oracle (not a test; prints a table)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vit as ovit          # noqa: E402
from oracle import cpu as ocpu          # noqa: E402
from transformer_explainability_b200 import _lib   # noqa: E402
from transformer_explainability_b200.baselines.ViT.ViT_LRP import VisionTransformer, vit_base_patch16_224  # noqa: E402


def rel(a, b):
    b = b.double()
    return ((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "vit_base_patch16_224"
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    print("usable cpus", ocpu.usable_cpus(), "cpu_count", os.cpu_count(), "threads", ocpu.set_torch_threads())
    params, heads = ovit.init_params(name, seed=0)
    xs = torch.randn(nb, 3, 224, 224, generator=torch.Generator().manual_seed(100))
    m = vit_base_patch16_224()
    m.load_state_dict(params)
    m = m.cuda().eval()
    eng = m.engine()
    maps, idx, logits = eng.explain(xs.cuda(), return_logits=True, flags=flags | _lib.FLAG_KEEP_ALL_CAMS)
    torch.cuda.synchronize()
    p64 = {k: v.double() for k, v in params.items()}
    for s in range(min(nb, 2)):
        t0 = time.time()
        ref, ridx, taps = ovit.explain(p64, xs[s:s + 1].double(), heads, return_taps=True)
        print("sample %d (oracle %.1fs) idx %d/%d logits rel %.2e  map rel %.3e abs %.3e" % (
            s, time.time() - t0, int(idx[s]), int(ridx), rel(logits[s], taps["logits"][0]), rel(maps[s], ref[0]),
            (maps[s].cpu().double() - ref[0]).abs().max().item()))
        for l in reversed(range(len(taps["grads"]))):
            ga = eng.tensor("attn", l)[s]
            gg = eng.tensor("attn_grad", l)[s]
            gc = eng.tensor("attn_cam", l)[s]
            mm = eng.tensor("rollout_mats", 0)[l, s]
            eye = torch.eye(mm.shape[0], dtype=torch.float64)
            print("  L%02d attn %.2e grad %.2e cam %.2e mat %.2e | sum(cam) gpu %.6f ref %.6f" % (
                l, rel(ga, taps["cache"]["blocks"][l]["attn"][0]), rel(gg, taps["grads"][l][0]),
                rel(gc, taps["cams"][l][0]), rel(mm.cpu().double() - eye, taps["mats"][l][0]),
                gc.double().sum().item(), taps["cams"][l][0].sum().item()))


if __name__ == "__main__":
    main()
