"""GPU diagnostic: every relprop rule at real ViT-B sizes (batch 4), each fed with the ORACLE's inputs for that
step, so that kernel error is separated from the chaotic amplification of the chain."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rules, vit as ovit, cpu as ocpu      # noqa: E402
from transformer_explainability_b200 import ops          # noqa: E402


def rel(a, b):
    b = b.double()
    return ((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def relb(a, b):      # per-sample
    return ["%.1e" % rel(a[i], b[i]) for i in range(a.shape[0])]


def g(t):
    return t.float().contiguous().cuda()


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    tc = len(sys.argv) > 2 and sys.argv[2] == "tc"
    ocpu.set_torch_threads()
    params, heads = ovit.init_params("vit_base_patch16_224", seed=0)
    xs = torch.randn(nb, 3, 224, 224, generator=torch.Generator().manual_seed(100))
    with torch.enable_grad():
        logits, cache = ovit.forward(params, xs, heads, need_grad=True)
    cache = {"cfg": cache["cfg"], "x_final_norm": cache["x_final_norm"].detach(),
             "blocks": [{k: v.detach() for k, v in c.items()} for c in cache["blocks"]]}
    idx = logits.argmax(-1)
    seed = torch.zeros_like(logits)
    seed[torch.arange(nb), idx] = 1
    p = params
    xf = cache["x_final_norm"]
    D = lambda t: t.double()
    r_cls = rules.linear_relprop(D(xf[:, 0]), D(p["head.weight"]), D(seed))
    r = rules.index_select_relprop(D(xf), r_cls.unsqueeze(1), 0)
    H = heads
    for i in (11, 10, 9):
        pre = "blocks.%d." % i
        c = cache["blocks"][i]
        r32 = r.float()
        # add2
        o1, o2 = rules.add_relprop(D(c["x_mid"]), D(c["mlp_out"]), D(r32))
        g1, g2 = ops.add_relprop(g(c["x_mid"]), g(c["mlp_out"]), g(r32))
        print("L%d add2   " % i, relb(g1, o1), relb(g2, o2))
        r1, r2 = o1.float(), o2.float()
        o = rules.linear_relprop(D(c["g"]), D(p[pre + "mlp.fc2.weight"]), D(r2))
        gg = ops.linear_relprop(g(c["g"]).reshape(-1, c["g"].shape[-1]), g(p[pre + "mlp.fc2.weight"]),
                                g(r2).reshape(-1, r2.shape[-1]), tensor_cores=tc).reshape(o.shape)
        print("L%d fc2    " % i, relb(gg, o))
        rf = o.float()
        o = rules.linear_relprop(D(c["xn2"]), D(p[pre + "mlp.fc1.weight"]), D(rf))
        gg = ops.linear_relprop(g(c["xn2"]).reshape(-1, 768), g(p[pre + "mlp.fc1.weight"]), g(rf).reshape(-1, rf.shape[-1]),
                                tensor_cores=tc).reshape(o.shape)
        print("L%d fc1    " % i, relb(gg, o))
        r2 = o.float()
        o = rules.clone_relprop(D(c["x_mid"]), (D(r1), D(r2)))
        gg = ops.clone_relprop(g(c["x_mid"]), [g(r1), g(r2)])
        print("L%d clone2 " % i, relb(gg, o))
        r32 = o.float()
        o1, o2 = rules.add_relprop(D(c["x_in"]), D(c["attn_out"]), D(r32))
        g1, g2 = ops.add_relprop(g(c["x_in"]), g(c["attn_out"]), g(r32))
        print("L%d add1   " % i, relb(g1, o1), relb(g2, o2))
        r1, r2 = o1.float(), o2.float()
        o = rules.linear_relprop(D(c["ctx"]), D(p[pre + "attn.proj.weight"]), D(r2))
        gg = ops.linear_relprop(g(c["ctx"]).reshape(-1, 768), g(p[pre + "attn.proj.weight"]), g(r2).reshape(-1, 768),
                                tensor_cores=tc).reshape(o.shape)
        print("L%d proj   " % i, relb(gg, o))
        rc = ovit._split_heads(o.float(), H)
        op, ov = rules.matmul_av_relprop(D(c["attn"]), D(c["v"]), D(rc))
        gp, gv = ops.matmul_av_relprop(g(c["attn"]), g(c["v"]), g(rc))
        print("L%d mm_av  " % i, relb(gp, op), relb(gv, ov))
        cam1, cam_v = (op / 2).float(), (ov / 2).float()
        oq, ok = rules.matmul_qk_relprop(D(c["q"]), D(c["k"]), D(cam1))
        gq, gk = ops.matmul_qk_relprop(g(c["q"]), g(c["k"]), g(cam1))
        print("L%d mm_qk  " % i, relb(gq, oq), relb(gk, ok))
        cq, ck = (oq / 2).float(), (ok / 2).float()
        rqkv = torch.cat([ovit._merge_heads(cq), ovit._merge_heads(ck), ovit._merge_heads(cam_v)], dim=-1)
        o = rules.linear_relprop(D(c["xn1"]), D(p[pre + "attn.qkv.weight"]), D(rqkv))
        gg = ops.linear_relprop(g(c["xn1"]).reshape(-1, 768), g(p[pre + "attn.qkv.weight"]), g(rqkv).reshape(-1, 2304),
                                tensor_cores=tc).reshape(o.shape)
        print("L%d qkv    " % i, relb(gg, o))
        r2 = o.float()
        o = rules.clone_relprop(D(c["x_in"]), (D(r1), D(r2)))
        gg = ops.clone_relprop(g(c["x_in"]), [g(r1), g(r2)])
        print("L%d clone1 " % i, relb(gg, o))
        r = o


if __name__ == "__main__":
    main()
