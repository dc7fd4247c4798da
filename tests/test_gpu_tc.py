"""GPU: the tcgen05 (TF32 tensor-core) z+ Linear rule against the fp32 SIMT path and the fp64 oracle.

Tolerance: TF32 operands carry a 10-bit mantissa (rna), accumulation is fp32, Z is a sum of non-negative
products -> relative error of a few 1e-4 on Z/S and on the output (stated: 2e-3 of the tensor maximum)."""
import pytest
import torch

from oracle import rules

pytestmark = pytest.mark.gpu


def rel(a, b):
    b = b.double()
    return ((a.double().cpu() - b.cpu()).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304),
                                           (256, 1024, 1024)])
def test_tc_linear_relprop_matches_simt_and_oracle(rows, inf, outf):
    from transformer_explainability_b200 import ops
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, inf, generator=g)
    w = torch.randn(outf, inf, generator=g) * 0.05
    r = torch.rand(rows, outf, generator=g)
    xd, wd, rd = x.cuda(), w.cuda(), r.cuda()
    simt = ops.linear_relprop(xd, wd, rd, tensor_cores=False)
    tc = ops.linear_relprop(xd, wd, rd, tensor_cores=True)
    torch.cuda.synchronize()
    ref = rules.linear_relprop(x.double(), w.double(), r.double())
    assert rel(simt, ref) < 2e-5
    assert rel(tc, ref) < 2e-3, "tcgen05 path: rel err %g" % rel(tc, ref)
    # conservation of relevance survives the reduced-precision operands
    assert abs(tc.double().sum().item() - r.double().sum().item()) < 2e-3 * r.sum().item()
    # single-pass variant fed with the saved forward output: Z = ((y - b) + |x||W|^T)/2 (what the engines run)
    b = torch.randn(outf, generator=g)
    y = ops.linear_forward(xd, wd, b.cuda())
    tc1 = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=b.cuda())
    assert rel(tc1, ref) < 3e-3, "single-pass tcgen05 path: rel err %g" % rel(tc1, ref)


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304)])
def test_tc_pair_kernels_cta_group2(rows, inf, outf):
    """The opt-in CTA-pair (tcgen05 cta_group::2) z+ kernels against the fp64 oracle and the single-CTA kernels
    (odd tile counts get an all-padding partner CTA)."""
    from transformer_explainability_b200 import _lib, ops
    g = torch.Generator().manual_seed(rows + 1)
    x = torch.randn(rows, inf, generator=g)
    w = torch.randn(outf, inf, generator=g) * 0.05
    r = torch.rand(rows, outf, generator=g)
    b = torch.randn(outf, generator=g)
    xd, wd, rd, bd = x.cuda(), w.cuda(), r.cuda(), b.cuda()
    y = ops.linear_forward(xd, wd, bd)
    one = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd)
    lib = _lib.load()
    _lib.check(lib.te_set_option(b"zplus_pair_kernels", 1), "te_set_option")
    try:
        pair = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd)
        pair_two_pass = ops.linear_relprop(xd, wd, rd, tensor_cores=True)          # two-pass S kernel + pair R kernel
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.te_set_option(b"zplus_pair_kernels", 0), "te_set_option")
    ref = rules.linear_relprop(x.double(), w.double(), r.double())
    assert rel(pair, ref) < 3e-3 and rel(pair_two_pass, ref) < 3e-3
    assert rel(pair, one.double()) < 1e-5          # same operands, same accumulation order per output tile


def test_tc_engine_vit_base_vs_simt_and_oracle():
    """ViT-B/16: engine with the tensor-core z+ path vs the fp32 SIMT engine and the fp64 oracle; medians over
    1e-7-perturbed copies (see tests/test_gpu_vit.py::_noise_trials for why)."""
    from oracle import cpu as ocpu
    from oracle import vit as ovit
    from transformer_explainability_b200 import _lib
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import vit_base_patch16_224
    from test_gpu_vit import _noise_trials, check_parity
    trials = 16
    params, heads = ovit.init_params("vit_base_patch16_224", seed=0)
    xs = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(100))
    m = vit_base_patch16_224()
    m.load_state_dict(params)
    m = m.cuda().eval()
    eng = m.engine()
    xb = torch.cat([_noise_trials(xs[s:s + 1], trials) for s in range(2)]).cuda()
    simt, idx0 = eng.explain(xb, flags=0)
    tc, idx1 = eng.explain(xb, flags=_lib.FLAG_ZPLUS_TENSOR_CORES)
    torch.cuda.synchronize()
    assert torch.equal(idx0, idx1)
    ocpu.set_torch_threads()
    for s in range(2):
        ref, ridx = ovit.explain({k: v.double() for k, v in params.items()}, xs[s:s + 1].double(), heads)
        check_parity(simt[s * trials:(s + 1) * trials], ref[0], "sample %d fp32 SIMT" % s)
        check_parity(tc[s * trials:(s + 1) * trials], ref[0], "sample %d tcgen05 z+" % s)


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304)])
def test_tc_3xtf32_linear_is_fp32_grade(rows, inf, outf):
    """Forward / backward Linear GEMMs on tcgen05 with the error-compensated 3xTF32 split.  The split removes the
    TF32 operand rounding (1e-3 -> 1e-6); what remains is the tensor core's own fp32 accumulation, which truncates
    (round-toward-zero) at every MMA, so the error grows linearly with the reduction length: measured 7e-9 * K
    (K = 3072: 2e-5) against 5e-7 for the fp32 SIMT kernel.  Stated bound: 1.5e-8 * K + 2e-6."""
    from transformer_explainability_b200 import ops
    g = torch.Generator().manual_seed(rows + 1)
    x = torch.randn(rows, inf, generator=g)
    w = torch.randn(outf, inf, generator=g) * 0.05
    b = torch.randn(outf, generator=g)
    dy = torch.randn(rows, outf, generator=g)
    ref_y = torch.nn.functional.linear(x.double(), w.double(), b.double())
    ref_dx = dy.double() @ w.double()
    for tc in (False, True):
        y = ops.linear_forward(x.cuda(), w.cuda(), b.cuda(), tensor_cores=tc)
        dx = ops.linear_backward(dy.cuda(), w.cuda(), tensor_cores=tc)
        torch.cuda.synchronize()
        ey, edx = rel(y, ref_y), rel(dx, ref_dx)
        print("rows %d in %d out %d tc=%s: fwd %.2e bwd %.2e" % (rows, inf, outf, tc, ey, edx))
        if tc:
            assert ey < 1.5e-8 * inf + 2e-6 and edx < 1.5e-8 * outf + 2e-6
        else:
            assert ey < 3e-6 and edx < 3e-6
    # the opt-in CTA-pair (tcgen05 cta_group::2) form of the same kernel: same operands, same chunked accumulation
    from transformer_explainability_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.te_set_option(b"linear_pair_kernels", 1), "te_set_option")
    try:
        y2 = ops.linear_forward(x.cuda(), w.cuda(), b.cuda(), tensor_cores=True)
        dx2 = ops.linear_backward(dy.cuda(), w.cuda(), tensor_cores=True)
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.te_set_option(b"linear_pair_kernels", 0), "te_set_option")
    assert rel(y2, y.double()) < 1e-6 and rel(dx2, dx.double()) < 1e-6


def test_tc_linear_engine_vit_base():
    """ViT-B/16 with every Linear GEMM (forward, backward, z+ rule) on tensor cores vs the fp32 SIMT engine and the
    fp64 oracle: class index bit-exact, logits at fp32 accuracy, maps in the same noise class (medians)."""
    from oracle import cpu as ocpu
    from oracle import vit as ovit
    from transformer_explainability_b200 import _lib
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import vit_base_patch16_224
    from test_gpu_vit import _noise_trials, check_parity
    trials = 16
    params, heads = ovit.init_params("vit_base_patch16_224", seed=0)
    xs = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(100))
    m = vit_base_patch16_224()
    m.load_state_dict(params)
    m = m.cuda().eval()
    eng = m.engine()
    xb = torch.cat([_noise_trials(xs[s:s + 1], trials) for s in range(2)]).cuda()
    simt, idx0, lg0 = eng.explain(xb, flags=0, return_logits=True)
    full = _lib.FLAG_TENSOR_CORES | _lib.FLAG_ROLLOUT_FUSED
    tc, idx1, lg1 = eng.explain(xb, flags=full, return_logits=True)
    torch.cuda.synchronize()
    assert torch.equal(idx0, idx1)
    ocpu.set_torch_threads()
    for s in range(2):
        ref, ridx, taps = ovit.explain({k: v.double() for k, v in params.items()}, xs[s:s + 1].double(), heads,
                                       return_taps=True)
        assert int(idx1[s * trials]) == int(ridx)
        assert rel(lg1[s * trials], taps["logits"][0]) < 1e-5
        check_parity(tc[s * trials:(s + 1) * trials], ref[0], "sample %d all-tensor-core" % s)


def test_tc_attention_contractions_engine():
    """Every attention-shaped contraction on tcgen05 (3xTF32): the N x N ones (QK^T, dctx V^T, attn_cam, S1; K-major
    operands) and the N x d ones reduced over tokens (attn v, attn^T dctx, dS k, dS^T q, S1 k, S1^T q, attn^T S2; MN-major
    tf32 operands in the SWIZZLE_128B_BASE32B layout).  Attention probabilities and attention gradients of every layer
    stay at fp32 accuracy (they chain through all of these kernels); maps in the same noise class."""
    from oracle import cpu as ocpu
    from oracle import vit as ovit
    from transformer_explainability_b200 import _lib
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import vit_base_patch16_224
    from test_gpu_vit import _noise_trials, check_parity
    trials = 16
    params, heads = ovit.init_params("vit_base_patch16_224", seed=0)
    xs = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(100))
    m = vit_base_patch16_224()
    m.load_state_dict(params)
    m = m.cuda().eval()
    eng = m.engine()
    ocpu.set_torch_threads()
    p64 = {k: v.double() for k, v in params.items()}
    ref, ridx, taps = ovit.explain(p64, xs[0:1].double(), heads, return_taps=True)
    maps, idx = eng.explain(xs[0:1].cuda(), flags=_lib.FLAG_ALL_FAST)
    assert int(idx[0]) == int(ridx)
    for l in (0, 6, 11):
        assert rel(m.blocks[l].attn.get_attn()[0], taps["cache"]["blocks"][l]["attn"][0]) < 1e-5
        assert rel(m.blocks[l].attn.get_attn_gradients()[0], taps["grads"][l][0]) < 1e-4
    assert rel(m.blocks[11].attn.get_attn_cam()[0], taps["cams"][11][0]) < 5e-2      # TF32 z+ rules feed this one
    xb = torch.cat([_noise_trials(xs[s:s + 1], trials) for s in range(2)]).cuda()
    fast, idx1 = eng.explain(xb, flags=_lib.FLAG_ALL_FAST)
    for s in range(2):
        r, _ = ovit.explain(p64, xs[s:s + 1].double(), heads)
        check_parity(fast[s * trials:(s + 1) * trials], r[0], "sample %d all-fast" % s)


@pytest.mark.parametrize("rows,inf,outf", [(394, 768, 3072), (1000, 3072, 768), (128, 256, 256)])
def test_tc_bf16_second_contraction(rows, inf, outf):
    """TE_FLAG_ZPLUS_BF16: S stored as bf16 and R_in = x+ (S W+) + x- (S W-) on tcgen05 kind::f16 with bf16 operands
    (8-bit mantissa, fp32 accumulate).  Stated tolerance 1.5e-2 of the tensor maximum; relevance is conserved to 1e-2."""
    from transformer_explainability_b200 import ops
    g = torch.Generator().manual_seed(rows + 7)
    x = torch.randn(rows, inf, generator=g)
    w = torch.randn(outf, inf, generator=g) * 0.05
    b = torch.randn(outf, generator=g)
    r = torch.rand(rows, outf, generator=g)
    xd, wd, rd, bd = x.cuda(), w.cuda(), r.cuda(), b.cuda()
    y = ops.linear_forward(xd, wd, bd)
    out = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd, bf16=True)
    torch.cuda.synchronize()
    ref = rules.linear_relprop(x.double(), w.double(), r.double())
    e = rel(out, ref)
    print("bf16 R kernel rows %d in %d out %d: rel %.2e" % (rows, inf, outf, e))
    assert e < 1.5e-2
    assert abs(out.double().sum().item() - r.double().sum().item()) < 1e-2 * r.sum().item()


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304),
                                           (50432, 768, 768)])
def test_tc_persistent_pair_kernels(rows, inf, outf):
    """The persistent CTA-pair (cta_group::2) kernels of te_tc_pair.cu — what the engines run by default: single-pass S
    kernel with the |x| transform + R kernel whose two products share one A tile, against the round-1 single-CTA kernels
    (same TF32 operands) and the fp64 oracle; the single-pass TF32 backward Linear against fp64 (TF32 operand error,
    stated 2e-3 of the tensor maximum).  Shapes include odd tile counts (all-padding partner CTA) and more tiles than
    clusters (several tiles per persistent cluster, both TMEM accumulator buffers reused)."""
    from transformer_explainability_b200 import _lib, ops
    g = torch.Generator().manual_seed(rows + 3)
    x = torch.randn(rows, inf, generator=g)
    w = torch.randn(outf, inf, generator=g) * 0.05
    r = torch.rand(rows, outf, generator=g)
    b = torch.randn(outf, generator=g)
    dy = torch.randn(rows, outf, generator=g)
    xd, wd, rd, bd = x.cuda(), w.cuda(), r.cuda(), b.cuda()
    y = ops.linear_forward(xd, wd, bd)
    new = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd)
    lib = _lib.load()
    _lib.check(lib.te_set_option(b"zplus_persistent", 0), "te_set_option")
    try:
        old = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd)
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.te_set_option(b"zplus_persistent", 1), "te_set_option")
    if rows <= 4096:
        ref = rules.linear_relprop(x.double(), w.double(), r.double())
        assert rel(new, ref) < 3e-3, "persistent pair z+ kernels: rel err %g" % rel(new, ref)
    assert rel(new, old.double()) < 2e-5          # same TF32 operands; only the accumulation order inside a tile differs
    assert abs(new.double().sum().item() - r.double().sum().item()) < 2e-3 * r.sum().item()
    dx = ops.linear_backward_tf32(dy.cuda(), wd)
    torch.cuda.synchronize()
    ref_dx = (dy.double().cuda() @ w.double().cuda()).cpu()
    e = rel(dx, ref_dx)
    print("tf32 pair backward rows %d in %d out %d: rel %.2e" % (rows, inf, outf, e))
    assert e < 2e-3


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304),
                                           (20000, 768, 768)])
def test_tc_mixed_kind_linear_is_fp32_grade(rows, inf, outf):
    """Forward Linear with the mixed-kind split (te_set_option("linear_mixed", 1)): main term TF32, the two correction terms
    as bf16 MMAs (SWIZZLE_64B operand tiles).  Same fp32-grade bound as the 3xTF32 kernel, every epilogue."""
    from transformer_explainability_b200 import _lib, ops
    g = torch.Generator().manual_seed(rows + 11)
    x = torch.randn(rows, inf, generator=g) * torch.logspace(-3, 1, inf)          # activations spanning four decades
    w = torch.randn(outf, inf, generator=g) * 0.05
    b = torch.randn(outf, generator=g)
    ref_y = torch.nn.functional.linear(x.double(), w.double(), b.double())
    lib = _lib.load()
    y3 = ops.linear_forward(x.cuda(), w.cuda(), b.cuda(), tensor_cores=True)
    try:
        _lib.check(lib.te_set_option(b"linear_mixed", 1), "te_set_option")
        ym = ops.linear_forward(x.cuda(), w.cuda(), b.cuda(), tensor_cores=True)
        _lib.check(lib.te_set_option(b"linear_mixed", 2), "te_set_option")         # persistent CTA-pair form
        yp = ops.linear_forward(x.cuda(), w.cuda(), b.cuda(), tensor_cores=True)
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.te_set_option(b"linear_mixed", 0), "te_set_option")
    e3, em, ep = rel(y3, ref_y), rel(ym, ref_y), rel(yp, ref_y)
    print("rows %d in %d out %d: 3xTF32 %.2e  mixed %.2e  mixed persistent pair %.2e" % (rows, inf, outf, e3, em, ep))
    assert em < 1.5e-8 * inf + 2e-6 and ep < 1.5e-8 * inf + 2e-6


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304)])
def test_tc_bf16_single_pass_denominator(rows, inf, outf):
    """TE_FLAG_ZPLUS_S1_BF16: the |x||W|^T term of the single-pass z+ denominator with bf16 operands (persistent pair kernel,
    kind::f16).  A sum of K non-negative products: the 2^-9 operand roundings average out; stated tolerance = the TF32 path's."""
    from transformer_explainability_b200 import ops
    g = torch.Generator().manual_seed(rows + 5)
    x = torch.randn(rows, inf, generator=g)
    w = torch.randn(outf, inf, generator=g) * 0.05
    b = torch.randn(outf, generator=g)
    r = torch.rand(rows, outf, generator=g)
    xd, wd, rd, bd = x.cuda(), w.cuda(), r.cuda(), b.cuda()
    y = ops.linear_forward(xd, wd, bd)
    tf = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd)
    bf = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd, bf16="s1")
    torch.cuda.synchronize()
    ref = rules.linear_relprop(x.double(), w.double(), r.double())
    e_tf, e_bf = rel(tf, ref), rel(bf, ref)
    print("bf16 S1 rows %d in %d out %d: TF32 %.2e  bf16 denominator %.2e" % (rows, inf, outf, e_tf, e_bf))
    assert e_bf < 3e-3
    assert abs(bf.double().sum().item() - r.double().sum().item()) < 2e-3 * r.sum().item()


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304),
                                           (20000, 768, 768), (300, 64, 256)])
def test_tc_f16_split_linear_is_fp32_grade(rows, inf, outf):
    """TE_FLAG_LINEAR_F16_SPLIT: forward Linear on tcgen05 kind::f16 with the row-scaled fp16 (hi, lo) split of both operands
    (te_tc_fwd16.cu).  fp16 carries the same 11-bit significand as TF32, so the bound is the 3xTF32 one; the per-row power-of-two
    scaling has to cope with rows and weight rows whose magnitudes span 12 decades, zero rows, and activations spanning four
    decades inside a row."""
    from transformer_explainability_b200 import ops
    g = torch.Generator().manual_seed(rows + 23)
    x = torch.randn(rows, inf, generator=g) * torch.logspace(-3, 1, inf)
    x = x * torch.logspace(-6, 6, rows)[:, None]                   # row magnitudes 1e-6 .. 1e6 (fp16 alone would over/underflow)
    x[rows // 2] = 0.0
    w = torch.randn(outf, inf, generator=g) * 0.05 * torch.logspace(-4, 2, outf)[:, None]
    w[3] = 0.0
    b = torch.randn(outf, generator=g)
    ref = torch.nn.functional.linear(x.double(), w.double())
    # without a bias: at row magnitudes of 1e-6 a bias of O(1) would swamp the product and hide the kernel's error
    y3 = ops.linear_forward(x.cuda(), w.cuda(), None, tensor_cores=True)
    yh = ops.linear_forward(x.cuda(), w.cuda(), None, tensor_cores=True, f16_split=True)
    yb = ops.linear_forward(x.cuda(), w.cuda(), b.cuda(), tensor_cores=True, f16_split=True)
    torch.cuda.synchronize()
    # error of every element relative to the scale of its own row and column
    scale = (x.double().abs() @ w.double().abs().T).clamp_min(1e-300) / inf ** 0.5
    e3 = ((y3.double().cpu() - ref).abs() / scale).max().item()
    eh = ((yh.double().cpu() - ref).abs() / scale).max().item()
    print("rows %d in %d out %d: 3xTF32 %.2e  fp16 split %.2e (per-element, relative to |x||W|^T / sqrt(K))" % (rows, inf, outf, e3, eh))
    assert torch.isfinite(yh).all()
    assert eh < 2e-5 and eh < 2 * e3 + 1e-6          # fp32 grade, and no worse than the 3xTF32 kernel on the same data
    assert (yh[rows // 2] == 0).all() and (yh[:, 3] == 0).all()                       # zero row / zero weight row: exactly zero
    assert torch.equal(yb.cpu(), (yh.cpu() + b))                                      # the bias is added last, in fp32


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304), (20000, 768, 768)])
def test_tc_fp16_second_contraction(rows, inf, outf):
    """TE_FLAG_ZPLUS_R_F16: R_in = x+ (S W+) + x- (S W-) on tcgen05 kind::f16 — S as hi-only block-scaled fp16 (one power of two
    per row and 128 columns), W+^T / W-^T as row-scaled fp16 (te_tc_fwd16.cu, FM_R).  Same 11 significant bits as the TF32 form
    (rounded to nearest): the rule error must not exceed the TF32 path's bound.  The relevance rows span 12 decades (S = R / Z
    inherits them): fp16 without the block scaling would over- / underflow."""
    from transformer_explainability_b200 import ops
    g = torch.Generator().manual_seed(rows + 31)
    x = torch.randn(rows, inf, generator=g)
    w = torch.randn(outf, inf, generator=g) * 0.05
    b = torch.randn(outf, generator=g)
    r = torch.rand(rows, outf, generator=g) * torch.logspace(-6, 6, rows)[:, None]
    r[rows // 3] = 0.0
    xd, wd, rd, bd = x.cuda(), w.cuda(), r.cuda(), b.cuda()
    y = ops.linear_forward(xd, wd, bd)
    tf = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd, bf16="s1")
    hf = ops.linear_relprop(xd, wd, rd, tensor_cores=True, y=y, bias=bd, bf16="s1", r_f16=True)
    torch.cuda.synchronize()
    ref = rules.linear_relprop(x.double(), w.double(), r.double())
    rowmax = ref.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    e_tf = ((tf.double().cpu() - ref).abs() / rowmax).max().item()        # per row: the rows differ by 12 decades
    e_hf = ((hf.double().cpu() - ref).abs() / rowmax).max().item()
    print("fp16 R rows %d in %d out %d: TF32 R %.2e  fp16 R %.2e (per row, relative to the row maximum)" % (rows, inf, outf, e_tf, e_hf))
    assert torch.isfinite(hf).all() and (hf[rows // 3] == 0).all()
    assert e_hf < 3e-3 and e_hf < 1.5 * e_tf + 1e-4
    rs, hs = r.double().sum(dim=1), hf.double().cpu().sum(dim=1)          # conservation per row (Z > 0 almost surely)
    assert ((hs - rs).abs() <= 3e-3 * rs.abs() + 1e-30).all()


@pytest.mark.parametrize("rows,inf,outf", [(128, 256, 256), (394, 768, 3072), (1000, 3072, 768), (77, 768, 2304), (20000, 768, 768)])
def test_tc_fp16_single_pass_backward(rows, inf, outf):
    """TE_FLAG_BACKWARD_F16: dx = dy W as ONE fp16 MMA per k-step (te_tc_fwd16.cu, FM_LIN1): block-scaled fp16 gradient rows that
    span 12 decades, row-scaled fp16 weights.  The operands keep TF32's 11 significant bits, rounded to nearest: the error
    stays below the single-pass TF32 kernel's bound, per row."""
    from transformer_explainability_b200 import ops
    g = torch.Generator().manual_seed(rows + 41)
    dy = torch.randn(rows, outf, generator=g) * torch.logspace(-9, 3, rows)[:, None]
    dy[rows // 3] = 0.0
    w = torch.randn(outf, inf, generator=g) * 0.05
    ref = dy.double() @ w.double()
    tf = ops.linear_backward_tf32(dy.cuda(), w.cuda())
    hf = ops.linear_backward_f16(dy.cuda(), w.cuda())
    torch.cuda.synchronize()
    rowmax = ref.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    e_tf = ((tf.double().cpu() - ref).abs() / rowmax).max().item()
    e_hf = ((hf.double().cpu() - ref).abs() / rowmax).max().item()
    print("fp16 backward rows %d in %d out %d: TF32 %.2e  fp16 %.2e (per row, relative to the row maximum)" % (rows, inf, outf, e_tf, e_hf))
    assert torch.isfinite(hf).all() and (hf[rows // 3] == 0).all()
    assert e_hf < 2e-3 and e_hf < 1.5 * e_tf + 1e-4


def test_f16_block_split_format_bit_exact():
    """The operand format of the fp16-split forward Linear, GPU pre-pass against its CPU restatement (oracle/f16_split.py): hi, lo and
    the block scales are BIT-EXACT (integer / byte work: exact power-of-two scaling, round-to-nearest-even fp16 conversions)."""
    import numpy as np
    from oracle import f16_split as F
    from transformer_explainability_b200 import ops
    g = torch.Generator().manual_seed(77)
    x = torch.randn(300, 768, generator=g) * torch.logspace(-3, 1, 768)
    x = x * torch.logspace(-30, 30, 300)[:, None]                # rows spanning 60 decades
    x[7] = 0.0
    x[9, 5] = float("inf")                                       # a non-finite block keeps scale 1 and propagates
    hi, lo, si = ops.f16_block_split(x.cuda())
    torch.cuda.synchronize()
    with np.errstate(invalid="ignore", over="ignore"):
        rh, rl, rs = F.split_rows(x.numpy())
    assert np.array_equal(si.cpu().numpy(), rs)
    assert np.array_equal(hi.cpu().numpy().view(np.uint16), rh.view(np.uint16))
    ok = np.ones_like(rh, dtype=bool)
    ok[9, :128] = False                                          # lo of the non-finite block is inf - inf = NaN (payload unspecified)
    assert np.array_equal(lo.cpu().numpy().view(np.uint16)[ok], rl.view(np.uint16)[ok])
