"""Closed-form CPU restatement of the reference's LRP rules (TEST INFRASTRUCTURE).

Each function states, without autograd, what the corresponding
``relprop`` of ``/root/reference/modules/layers_ours.py`` computes for
``alpha=1`` (the only value any caller passes:
``baselines/ViT/ViT_explanation_generator.py:27``,
``BERT_explainability/modules/BERT/ExplanationGenerator.py:31``).

All reductions that the reference performs over a whole ``B=1`` tensor are
performed here PER SAMPLE (leading dim), because the reference path is only
correct at ``B=1`` (``ViT_LRP.py:362-363``, ``layers_ours.py:109-116``):
a batch here means "independent B=1 explanations".

Works in any float dtype (fp32 mirrors the reference; fp64 is the yard-stick).
"""
import torch

EPS = 1e-9


def safe_divide(a, b):
    """``layers_ours.py:10-13``.

    den = clamp(b,min=eps)+clamp(b,max=eps) == b+eps ; den += eps where den==0 ;
    result = a/den * (b != 0).
    """
    den = b.clamp(min=EPS) + b.clamp(max=EPS)
    den = den + den.eq(0).to(den.dtype) * EPS
    return a / den * b.ne(0).to(b.dtype)


def linear_relprop(x, w, r):
    """z+ rule, ``Linear.relprop`` (``layers_ours.py:207-230``) with alpha=1, beta=0.

    x [..., in], w [out, in], r [..., out] -> [..., in].
    Z = x+ W+^T + x- W-^T ; S = sd(R, Z) ; R_in = x+ * (S W+) + x- * (S W-).
    The inhibitor half is multiplied by beta = 0 in the reference and dropped.
    """
    pw = w.clamp(min=0)
    nw = w.clamp(max=0)
    px = x.clamp(min=0)
    nx = x.clamp(max=0)
    z = px @ pw.t() + nx @ nw.t()
    s = safe_divide(r, z)
    return px * (s @ pw) + nx * (s @ nw)


def linear_relprop_lrp(x, w, r):
    """``Linear.relprop`` of the ``layers_lrp`` baseline variant (``modules/layers_lrp.py:187-210``, alpha=1): the two
    halves are divided by their OWN denominators, S1 = sd(R, x+ W+^T), S2 = sd(R, x- W-^T) (``:199-200``; the
    ``layers_ours`` rule divides both by the sum), R_in = x+ * (S1 W+) + x- * (S2 W-)."""
    pw = w.clamp(min=0)
    nw = w.clamp(max=0)
    px = x.clamp(min=0)
    nx = x.clamp(max=0)
    s1 = safe_divide(r, px @ pw.t())
    s2 = safe_divide(r, nx @ nw.t())
    return px * (s1 @ pw) + nx * (s2 @ nw)


def add_relprop_simple(x1, x2, r):
    """``Add.relprop`` of the ``layers_lrp`` variant = ``RelPropSimple.relprop`` (``modules/layers_lrp.py:48-60,98-100``):
    S = sd(R, x1 + x2) ; outputs x1 * S, x2 * S — no ratio re-normalisation."""
    s = safe_divide(r, x1 + x2)
    return x1 * s, x2 * s


def _per_sample_sum(t):
    return t.reshape(t.shape[0], -1).sum(dim=1).reshape([-1] + [1] * (t.dim() - 1))


def add_relprop(x1, x2, r):
    """``Add.relprop`` (``layers_ours.py:97-120``), reductions per sample.

    x2 may broadcast against x1 (BERT's additive mask [B,1,1,S]); its relevance
    is then sum-reduced over the broadcast dims like autograd does.
    """
    z = x1 + x2
    s = safe_divide(r, z)
    a = x1 * s
    b_full = x2 * s                     # x2 * dZ/dx2 before the broadcast reduction
    if x2.shape != x1.shape:
        red = [d for d in range(x1.dim()) if x2.shape[d] == 1 and x1.shape[d] != 1]
        b = x2 * s.sum(dim=red, keepdim=True)
    else:
        b = b_full
    a_sum = _per_sample_sum(a)
    b_sum = _per_sample_sum(b)
    r_sum = _per_sample_sum(r)
    den = a_sum.abs() + b_sum.abs()
    a_fact = safe_divide(a_sum.abs(), den) * r_sum
    b_fact = safe_divide(b_sum.abs(), den) * r_sum
    a = a * safe_divide(a_fact, a_sum)
    b = b * safe_divide(b_fact, b_sum)
    return a, b


def clone_relprop(x, rs):
    """``Clone.relprop`` (``layers_ours.py:151-169``): R = X * sum_i sd(R_i, X)."""
    c = None
    for r in rs:
        t = safe_divide(r, x)
        c = t if c is None else c + t
    return x * c


def index_select_relprop(x, r, index=0):
    """``IndexSelect.relprop`` (``layers_ours.py:129-147``) for dim=1, one index.

    x [B,N,D], r [B,1,D] -> [B,N,D], non-zero only at token ``index``.
    """
    out = torch.zeros_like(x)
    xi = x[:, index:index + 1, :]
    out[:, index:index + 1, :] = xi * safe_divide(r, xi)
    return out


def matmul_av_relprop(p, v, r):
    """matmul2 rule: ``einsum('bhij,bhjd->bhid')`` through ``RelPropSimple.relprop``
    (``layers_ours.py:48-60,122-127``).  Returns UN-halved (R_P, R_V); the caller
    halves both (``ViT_LRP.py:161-162``)."""
    z = p @ v
    s = safe_divide(r, z)
    return p * (s @ v.transpose(-1, -2)), v * (p.transpose(-1, -2) @ s)


def matmul_qk_relprop(q, k, r):
    """matmul1 rule: ``einsum('bhid,bhjd->bhij')`` (unscaled Z = Q K^T,
    ``ViT_LRP.py:139``).  Returns UN-halved (R_Q, R_K)."""
    z = q @ k.transpose(-1, -2)
    s = safe_divide(r, z)
    return q * (s @ k), k * (s.transpose(-1, -2) @ q)


def conv_zb_relprop(x, w, r, stride):
    """z^B ("box") rule of the first layer, ``Conv2d.relprop`` for a 3-channel input
    (``layers_ours.py:242-259``), per sample.

    x [B,3,H,W] image, w [O,3,P,P], r [B,O,H/P,W/P] -> [B,3,H,W].
    L / H = per-sample min / max of the image (broadcast to the image shape);
    Za = conv(x, W) - conv(L, W+) - conv(H, W-) + 1e-9 ; S = R / Za (plain division) ;
    C = x * convT(S, W) - L * convT(S, W+) - H * convT(S, W-).
    """
    import torch.nn.functional as F
    pw = w.clamp(min=0)
    nw = w.clamp(max=0)
    lo = x * 0 + x.amin(dim=(1, 2, 3), keepdim=True)
    hi = x * 0 + x.amax(dim=(1, 2, 3), keepdim=True)
    za = F.conv2d(x, w, None, stride=stride) - F.conv2d(lo, pw, None, stride=stride) - \
        F.conv2d(hi, nw, None, stride=stride) + 1e-9
    s = r / za

    def back(wt):
        return F.conv_transpose2d(s, wt, stride=stride)

    return x * back(w) - lo * back(pw) - hi * back(nw)


def aggregate(grad, cam):
    """``ViT_LRP.py:359-365`` / ``ExplanationGenerator.py:49-55``:
    grad, cam [B,H,N,N] -> mean_h relu(grad*cam) [B,N,N] (clamp THEN mean)."""
    return (grad * cam).clamp(min=0).mean(dim=1)


def rollout(mats, start_layer=0, normalize=False):
    """``compute_rollout_attention``: ViT (``ViT_LRP.py:38-49``, normalize=False)
    and BERT (``ExplanationGenerator.py:7-18``, normalize=True).
    mats: list of [B,N,N] -> [B,N,N]."""
    n = mats[0].shape[1]
    eye = torch.eye(n, dtype=mats[0].dtype).expand_as(mats[0])
    ms = [m + eye for m in mats]
    if normalize:
        ms = [m / m.sum(dim=-1, keepdim=True) for m in ms]
    joint = ms[start_layer]
    for i in range(start_layer + 1, len(ms)):
        joint = ms[i].bmm(joint)
    return joint
