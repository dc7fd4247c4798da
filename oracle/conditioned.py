"""Well-conditioned full-size parameter sets (TEST INFRASTRUCTURE).

At random init the reference's rules are ill-conditioned at ViT-B / BERT-base size: ``Add.relprop``
(``modules/layers_ours.py:97-120``) divides by the signed element-wise sum ``x1 + x2`` and both attention matmul rules
(``:48-60``) divide by signed ``Q K^T`` / ``A V``; with ~10^5-10^7 elements per tensor some denominators are ~1e-7 and
the reference's own fp32 result differs from its fp64 result by 1e-5 ... 7e-1 depending on the input (SURVEY.md §8c,
``tests/test_gpu_vit.py::_noise_trials``).  That noise hides kernel errors at full size.

``condition_vit`` / ``condition_bert`` shift a random-init ``state_dict`` so that every such denominator is bounded
away from zero: positive offsets on the q / k / v biases (``Q K^T > 0``, ``A V > 0``), on the biases of the two
residual branches and on the token stream (``x1 + x2 > 0``).  Signed LayerNorm outputs and signed weights still
exercise both halves of the z+ rule; relevance stays non-negative above block 0, so nothing cancels.  In this regime
the fp32 oracle agrees with the fp64 oracle to ~1e-5 of the map maximum, so a GPU-vs-fp64 comparison with a tight
relative bound is a test that can fail.
"""
import torch


def condition_vit(params, c_qkv=4.0, c_proj=8.0, c_fc2=4.0, c_tok=4.0):
    p = {k: v.clone() for k, v in params.items()}
    depth = 1 + max(int(k.split(".")[1]) for k in p if k.startswith("blocks."))
    for i in range(depth):
        pre = "blocks.%d." % i
        p[pre + "attn.qkv.bias"] = p[pre + "attn.qkv.bias"] + c_qkv
        p[pre + "attn.proj.bias"] = p[pre + "attn.proj.bias"] + c_proj
        p[pre + "mlp.fc2.bias"] = p[pre + "mlp.fc2.bias"] + c_fc2
    p["pos_embed"] = p["pos_embed"] + c_tok
    return p


def condition_bert(params, c_qkv=3.0, c_dense=16.0, c_ln=0.0):
    p = {k: v.clone() for k, v in params.items()}
    for k in list(p):
        if k.endswith("attention.self.query.bias") or k.endswith("attention.self.key.bias") or \
                k.endswith("attention.self.value.bias"):
            p[k] = p[k] + c_qkv
        elif k.endswith("attention.output.dense.bias") or (k.endswith("output.dense.bias") and "attention" not in k):
            p[k] = p[k] + c_dense
        elif k.endswith("LayerNorm.bias"):
            p[k] = p[k] + c_ln
    return p
