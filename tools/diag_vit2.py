"""GPU diagnostic: stop the engine relprop at block `stop` (start_layer=stop) and compare the scratch tensors it
leaves behind with the oracle's intermediates of that block."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vit as ovit, cpu as ocpu          # noqa: E402
from transformer_explainability_b200.baselines.ViT.ViT_LRP import vit_base_patch16_224  # noqa: E402


def rel(a, b):
    b = b.double()
    return ((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    stops = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [10]
    ocpu.set_torch_threads()
    params, heads = ovit.init_params("vit_base_patch16_224", seed=0)
    xs = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(100))[first:first + nb]
    m = vit_base_patch16_224()
    m.load_state_dict(params)
    m = m.cuda().eval()
    eng = m.engine()
    p64 = {k: v.double() for k, v in params.items()}
    ref = {}
    for s in range(nb):
        out, idx, taps = ovit.explain(p64, xs[s:s + 1].double(), heads, return_taps=True)
        ref[s] = taps
    for stop in stops:
        eng.forward(xs.cuda())
        maps, idx = eng.attribute(start_layer=stop)
        torch.cuda.synchronize()
        print("== stop at block", stop)
        for s in range(nb):
            t = ref[s]["relprop"][stop]
            row = []
            row.append("clone2(d0) %.1e" % rel(eng.tensor("tmp_d0")[s], t["clone2"][0]))
            row.append("add1_r1(d1) %.1e" % rel(eng.tensor("tmp_d1")[s], t["add1_r1"][0]))
            row.append("add1_r2(d2) %.1e" % rel(eng.tensor("tmp_d2")[s], t["add1_r2"][0]))
            row.append("proj(d3) %.1e" % rel(eng.tensor("tmp_d3")[s], t["proj"][0]))
            row.append("fc2(f0) %.1e" % rel(eng.tensor("tmp_f0")[s], t["fc2"][0]))
            row.append("cam %.1e" % rel(eng.tensor("attn_cam", stop)[s], ref[s]["cams"][stop][0]))
            if stop + 1 in ref[s]["relprop"]:
                tp = ref[s]["relprop"][stop + 1]
                row.append("| prev r_qkv(3d1) %.1e" % rel(eng.tensor("tmp_3d1")[s], tp["r_qkv"][0]))
                row.append("prev cam %.1e" % rel(eng.tensor("attn_cam", stop + 1)[s], ref[s]["cams"][stop + 1][0]))
            print(" s%d" % s, "  ".join(row))


if __name__ == "__main__":
    main()
