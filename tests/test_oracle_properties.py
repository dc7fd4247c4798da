"""CPU: size-independent properties of the oracle's rules (hypothesis-driven shapes / seeds) — the invariants the
GPU parity tests rely on at full batch size: relevance conservation (the authors' own check, ViT_LRP.py:326,334),
per-sample independence of a batch, and the algebra of the rollout."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import rules

SET = dict(max_examples=25, deadline=None, derandomize=True)      # fixed example set: the suite must be reproducible


def _t(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


@settings(**SET)
@given(st.integers(0, 10_000), st.integers(1, 4), st.integers(1, 9), st.integers(1, 24), st.integers(1, 24))
def test_linear_rule_conserves_relevance(seed, b, n, fin, fout):
    x, w = _t(seed, b, n, fin), _t(seed + 1, fout, fin, scale=0.3)
    r = _t(seed + 2, b, n, fout).abs()
    out = rules.linear_relprop(x, w, r)
    z = x.clamp(min=0) @ w.clamp(min=0).t() + x.clamp(max=0) @ w.clamp(max=0).t()
    # safe_divide: R * Z / (Z + 1e-9), and relevance routed through an exactly-zero denominator is dropped
    kept = (r * z / (z + 1e-9) * z.ne(0)).sum()
    assert torch.allclose(out.sum(), kept, rtol=1e-9, atol=1e-12)


@settings(**SET)
@given(st.integers(0, 10_000), st.integers(1, 4), st.integers(2, 40))
def test_add_rule_conserves_relevance_per_sample_and_is_batch_independent(seed, b, n):
    x1, x2, r = _t(seed, b, n), _t(seed + 1, b, n), _t(seed + 2, b, n)
    a, c = rules.add_relprop(x1, x2, r)
    total = a.reshape(b, -1).sum(1) + c.reshape(b, -1).sum(1)
    assert torch.allclose(total, r.reshape(b, -1).sum(1), rtol=1e-7, atol=1e-9)
    for s in range(b):                                # a batch is a set of independent B=1 explanations
        a1, c1 = rules.add_relprop(x1[s:s + 1], x2[s:s + 1], r[s:s + 1])
        assert torch.allclose(a1[0], a[s], rtol=1e-12, atol=0) and torch.allclose(c1[0], c[s], rtol=1e-12, atol=0)


@settings(**SET)
@given(st.integers(0, 10_000), st.integers(1, 3), st.integers(2, 30))
def test_clone_rule_inverts_a_proportional_split(seed, b, n):
    x = _t(seed, b, n)
    s1, s2 = _t(seed + 1, b, n), _t(seed + 2, b, n)
    out = rules.clone_relprop(x, (x * s1, x * s2))   # R_i = X * s_i  ->  X * sum_i sd(R_i, X) ~= X * (s1 + s2)
    damp = x / (x + 1e-9)                              # safe_divide's epsilon
    assert torch.allclose(out, x * (s1 + s2) * damp, rtol=1e-9, atol=1e-12)


@settings(**SET)
@given(st.integers(0, 10_000), st.integers(1, 3), st.integers(2, 5), st.integers(2, 12), st.booleans())
def test_rollout_algebra(seed, b, layers, n, normalize):
    mats = [_t(seed + i, b, n, n).abs() * 0.1 for i in range(layers)]
    eye = torch.eye(n, dtype=torch.float64)
    aug = [m + eye for m in mats]
    if normalize:
        aug = [m / m.sum(dim=-1, keepdim=True) for m in aug]
    last = rules.rollout(mats, start_layer=layers - 1, normalize=normalize)
    assert torch.allclose(last, aug[-1])
    full = rules.rollout(mats, start_layer=0, normalize=normalize)
    tail = rules.rollout(mats, start_layer=1, normalize=normalize)
    assert torch.allclose(full, tail.bmm(aug[0]), rtol=1e-10, atol=1e-12)      # J_0 = J_1 (M_0 + I)
    if normalize:
        assert torch.allclose(full.sum(dim=-1), torch.ones(b, n, dtype=torch.float64))   # rows stay stochastic
    # the row-only form the fused kernel computes: e_0^T M_{L-1} ... M_s
    row = torch.zeros(b, 1, n, dtype=torch.float64)
    row[:, 0, 0] = 1
    for m in reversed(aug):
        row = row.bmm(m)
    assert torch.allclose(row[:, 0], full[:, 0], rtol=1e-10, atol=1e-12)


@settings(**SET)
@given(st.integers(0, 10_000), st.integers(1, 2), st.integers(1, 3), st.integers(2, 10), st.integers(1, 8))
def test_matmul_rules_conserve_relevance(seed, b, h, n, d):
    p = _t(seed, b, h, n, n).softmax(-1)
    v = _t(seed + 1, b, h, n, d)
    r = _t(seed + 2, b, h, n, d)
    rp, rv = rules.matmul_av_relprop(p, v, r)
    z = p @ v
    kept = (r * z / (z + 1e-9) * z.ne(0)).sum()       # safe_divide's epsilon; each operand receives all of R
    assert torch.allclose(rp.sum(), kept, rtol=1e-9, atol=1e-12) and torch.allclose(rv.sum(), kept, rtol=1e-9, atol=1e-12)
