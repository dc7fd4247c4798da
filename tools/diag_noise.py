"""GPU diagnostic: is a large engine-vs-oracle deviation on one sample a rare chaotic event (disappears under
1e-7 input perturbations) or systematic?  Prints map error vs the fp64 oracle of the perturbed input."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vit as ovit, cpu as ocpu          # noqa: E402
from transformer_explainability_b200.baselines.ViT.ViT_LRP import vit_base_patch16_224  # noqa: E402


def main():
    sample = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ocpu.set_torch_threads()
    params, heads = ovit.init_params("vit_base_patch16_224", seed=0)
    x0 = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(100))[sample:sample + 1]
    m = vit_base_patch16_224()
    m.load_state_dict(params)
    m = m.cuda().eval()
    eng = m.engine()
    p64 = {k: v.double() for k, v in params.items()}
    xs = [x0] + [x0 * (1 + 1e-7 * torch.randn(x0.shape, generator=torch.Generator().manual_seed(k))) for k in range(1, trials)]
    xb = torch.cat(xs)
    maps, idx = eng.explain(xb.cuda(), flags=flags)
    ref0, _ = ovit.explain(p64, x0.double(), heads)
    o32, _ = ovit.explain(params, x0, heads)
    sc = ref0.abs().max().item()
    print("oracle fp32 vs fp64 on unperturbed input: %.3e" % ((o32.double() - ref0).abs().max().item() / sc))
    errs = sorted((maps[k].cpu().double() - ref0[0]).abs().max().item() / sc for k in range(trials))
    print("sample %d flags %d: map rel err vs fp64 oracle over %d trials (sorted): %s  median %.2e" % (
        sample, flags, trials, " ".join("%.1e" % e for e in errs), 0.5 * (errs[trials // 2 - 1] + errs[trials // 2])))


if __name__ == "__main__":
    main()
