"""Drop-in for ``baselines/ViT/ViT_orig_LRP.py``: the same ViT on the rule library of ``modules/layers_lrp.py``
(Linear with separate denominators ``:199-200``, Add without ratio normalisation ``:98-100``) — the "LRP" / "partial LRP"
baselines of ``generate_visualizations.py:79,84`` and ``imagenet_seg_eval.py``.  Same engine, ``TE_FLAG_RULES_LRP``;
``relprop``'s default ``method`` is ``"grad"`` like the reference's (``ViT_orig_LRP.py:324``), and
``"transformer_attribution"`` is not a name this variant knows (``:355``)."""
from transformer_explainability_b200 import _lib
from . import ViT_LRP as _base
from .ViT_LRP import compute_rollout_attention, _conv_filter        # noqa: F401  (same module-level names as the reference)


class VisionTransformer(_base.VisionTransformer):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._rule_flags = _lib.FLAG_RULES_LRP

    def relprop(self, cam=None, method="grad", is_ablation=False, start_layer=0, **kwargs):
        if method == "transformer_attribution":
            raise ValueError("ViT_orig_LRP has no method 'transformer_attribution' (ViT_orig_LRP.py:355): use 'grad'")
        return super().relprop(cam, method=method, is_ablation=is_ablation, start_layer=start_layer, **kwargs)


def vit_base_patch16_224(pretrained=False, **kwargs):
    _base._no_pretrained(pretrained)
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True, **kwargs)


def vit_large_patch16_224(pretrained=False, **kwargs):
    _base._no_pretrained(pretrained)
    return VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, qkv_bias=True, **kwargs)
