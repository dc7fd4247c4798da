"""GPU: the ViT engine through the C ABI / the reference-shaped facade vs the oracle and the golden fixtures.

Tolerances (fp32 engine vs fp64 oracle):
 * argmax class index: bit-exact;
 * logits, attention, attention gradients: relative 1e-4 of the tensor maximum;
 * relevance maps: absolute L_inf <= 1e-4 on raw maps (BASELINE.json north_star), and — because the reference
   itself is ill-conditioned at random init — relative error bounded by a multiple of the
   reference's own fp32-vs-fp64 error where that is stored, or by 2e-2 on the well-conditioned tiny model.
"""
import os

import numpy as np
import pytest
import torch

from oracle import vit as ovit

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel(a, b):
    b = torch.as_tensor(b).double()
    return ((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def make_model(params, heads, **kw):
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import VisionTransformer
    m = VisionTransformer(qkv_bias=True, num_heads=heads, **kw)
    m.load_state_dict({k: v.float() for k, v in params.items()})
    return m.cuda().eval()


TINY = dict(img_size=32, patch_size=8, embed_dim=64, depth=3, mlp_ratio=4., num_classes=10)


def test_tiny_vs_golden_reference(golden_dir):
    """Per-block taps and final maps vs the UNMODIFIED reference's outputs (fp64 run) — well-conditioned model."""
    from transformer_explainability_b200.baselines.ViT.ViT_explanation_generator import LRP
    g = np.load(os.path.join(golden_dir, "vit_tiny.npz"))
    params = {k[len("param."):]: T(g[k]) for k in g.files if k.startswith("param.")}
    model = make_model(params, int(g["heads"]), **TINY)
    lrp = LRP(model)
    x = T(g["x"]).cuda()
    for s in range(x.shape[0]):
        for sl in (0, 1):
            out = lrp.generate_LRP(x[s:s + 1], start_layer=sl)
            ref = T(g["f64.s%d.map.sl%d" % (s, sl)])
            assert out.shape == ref.shape == (1, 16)
            assert (out.cpu().double() - ref).abs().max() < 1e-4
            assert rel(out, ref) < 2e-2, "tiny map s=%d sl=%d rel=%g" % (s, sl, rel(out, ref))
        out = lrp.generate_LRP(x[s:s + 1], start_layer=0)
        logits = model._engine.tensor("logits")
        assert rel(logits, T(g["f64.s%d.logits" % s])) < 1e-5
        assert int(logits.argmax()) == int(T(g["f64.s%d.logits" % s]).argmax())
        for l in range(3):
            assert rel(model.blocks[l].attn.get_attn(), T(g["f64.s%d.attn.%d" % (s, l)])) < 1e-5
            assert rel(model.blocks[l].attn.get_attn_gradients(), T(g["f64.s%d.grad.%d" % (s, l)])) < 1e-4
            assert rel(model.blocks[l].attn.get_attn_cam(), T(g["f64.s%d.cam.%d" % (s, l)])) < 2e-2
    out = lrp.generate_LRP(x[0:1], index=3)
    assert rel(out, T(g["f64.s0.map.index3"])) < 2e-2


def test_tiny_batched_equals_single(golden_dir):
    """Batch = independent B=1 explanations: the batched call reproduces the per-sample calls."""
    from transformer_explainability_b200.baselines.ViT.ViT_explanation_generator import LRP
    params, heads = ovit.init_params("vit_tiny_test", seed=3, rand_affine=True)
    model = make_model(params, heads, **TINY)
    x = torch.randn(5, 3, 32, 32, generator=torch.Generator().manual_seed(0)).cuda()
    lrp = LRP(model)
    batched, idx = lrp.generate_LRP_batched(x, return_index=True)
    ref, ridx = ovit.explain({k: v.double() for k, v in params.items()}, x.cpu().double(), heads)
    assert torch.equal(idx.cpu().long(), ridx)
    assert rel(batched, ref) < 2e-2
    for s in range(5):
        one = lrp.generate_LRP(x[s:s + 1])
        assert torch.allclose(one[0], batched[s], rtol=1e-5, atol=1e-9)
    # chunked execution gives the same answer
    chunked = lrp.generate_LRP_batched(x, chunk=2)
    assert torch.allclose(chunked, batched, rtol=1e-5, atol=1e-9)


def test_distilled_extension_vs_oracle():
    params, heads = ovit.init_params("vit_tiny_test", seed=4, rand_affine=True, distilled=True)
    model = make_model(params, heads, distilled=True, **TINY)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    maps, idx = model.engine().explain(x.cuda())
    ref, ridx = ovit.explain({k: v.double() for k, v in params.items()}, x.double(), heads)
    assert maps.shape == (2, 16)
    assert torch.equal(idx.cpu().long(), ridx)
    assert rel(maps, ref) < 2e-2


def _base_case(name, n, seed_w=0, seed_x=100):
    params, heads = ovit.init_params(name, seed=seed_w)
    xs = torch.randn(n, 3, 224, 224, generator=torch.Generator().manual_seed(seed_x))
    return params, heads, xs


@pytest.mark.parametrize("name", ["vit_base_patch16_224"])
def test_vit_base_vs_oracle_and_golden(golden_dir, name):
    """ViT-B/16 (BASELINE configs[0]/[1] shape): engine vs fp64 oracle run on this box, and vs the reference's
    stored maps.  err_new is judged against err_ref = |ref32 - ref64|."""
    g = np.load(os.path.join(golden_dir, "vit_base.npz"))
    n = int(g["n"])
    params, heads, xs = _base_case(name, n, int(g["param_seed"]), int(g["x_seed"]))
    model = make_model(params, heads)
    maps, idx, logits = model.engine().explain(xs.cuda(), return_logits=True)
    torch.cuda.synchronize()
    assert maps.shape == (n, 196)
    p64 = {k: v.double() for k, v in params.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    for s in range(2):
        ref, ridx, taps = ovit.explain(p64, xs[s:s + 1].double(), heads, return_taps=True)
        assert int(idx[s]) == int(ridx)                                   # bit-exact class index
        assert rel(logits[s], taps["logits"][0]) < 1e-4
        err = (maps[s].cpu().double() - ref[0]).abs().max().item()
        assert err <= 1e-4, "raw map L_inf %g" % err
        for l in (0, 5, 11):
            assert rel(model.blocks[l].attn.get_attn_gradients()[s], taps["grads"][l][0]) < 1e-3
    reproducible = abs(sum(v.double().sum().item() for v in params.values()) - float(g["w_checksum"])) < 1e-6 * abs(
        float(g["w_checksum"]))
    if reproducible:
        for s in range(n):
            assert int(idx[s]) == int(g["f64.index"][s]) == int(g["f32.index"][s])
            ref64, ref32 = g["f64.maps"][s], g["f32.maps"][s]
            err_ref = np.abs(ref32 - ref64).max()
            err_new = np.abs(maps[s].cpu().numpy().astype(np.float64) - ref64).max()
            assert err_new <= max(1e-4, 20 * err_ref)
            scale = np.abs(ref64).max()
            print("sample %d: err_new/max %.3g  err_ref/max %.3g" % (s, err_new / scale, err_ref / scale))


def test_vit_large_smoke_vs_oracle():
    """ViT-L/16 (BASELINE configs[2]): one sample vs the fp64 oracle."""
    params, heads, xs = _base_case("vit_large_patch16_224", 1, 1, 7)
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import vit_large_patch16_224
    m = vit_large_patch16_224()
    m.load_state_dict(params)
    m = m.cuda().eval()
    maps, idx = m.engine().explain(xs.cuda())
    ref, ridx = ovit.explain({k: v.double() for k, v in params.items()}, xs.double(), heads)
    assert int(idx[0]) == int(ridx)
    assert (maps.cpu().double() - ref).abs().max() <= 1e-4


def test_full_batch_properties():
    """BASELINE batch size through the public call: finite, non-negative row-0 rollout, permutation-equivariant."""
    params, heads, xs = _base_case("vit_base_patch16_224", 8, 0, 5)
    model = make_model(params, heads)
    eng = model.engine()
    xb = xs.repeat(4, 1, 1, 1)[:32].cuda()
    maps, idx = eng.explain(xb)
    assert torch.isfinite(maps).all()
    assert (maps >= 0).all()                      # products of (relu-mean + I) matrices are non-negative
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(0)).cuda()
    maps_p, idx_p = eng.explain(xb[perm])
    assert torch.equal(idx_p, idx[perm])
    assert torch.allclose(maps_p, maps[perm], rtol=1e-5, atol=1e-10)
