"""CPU: the C-ABI library loads and exports exactly what include/te_b200.h declares; host-side tables agree."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "te_b200.h")).read()
    return sorted(set(re.findall(r"TE_API[^;(]*?\b(te_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from transformer_explainability_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "library does not export %s" % s
    assert set(syms) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.te_version() >= 100


def test_weight_table_matches_facade_state_dict():
    from transformer_explainability_b200 import _lib
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import VisionTransformer
    lib = _lib.load()
    for distilled in (False, True):
        m = VisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=2, num_heads=4, num_classes=10,
                              qkv_bias=True, distilled=distilled)
        cfg = m._cfg
        n = lib.te_vit_num_weights(ctypes.byref(cfg))
        sd = m.state_dict()
        seen, end = set(), 0
        for i in range(n):
            name = lib.te_vit_weight_name(ctypes.byref(cfg), i).decode()
            numel = lib.te_vit_weight_numel(ctypes.byref(cfg), i)
            off = lib.te_vit_weight_offset(ctypes.byref(cfg), i)
            assert sd[name].numel() == numel
            assert off % 32 == 0 and off >= end
            end = off + numel
            seen.add(name)
        assert seen == set(sd.keys())
        assert lib.te_vit_weight_total(ctypes.byref(cfg)) >= end


def test_workspace_query_and_errors_without_gpu():
    from transformer_explainability_b200 import _lib
    from transformer_explainability_b200.engine import vit_config
    lib = _lib.load()
    cfg = vit_config()
    b1 = lib.te_vit_workspace_bytes(ctypes.byref(cfg), 1)
    b2 = lib.te_vit_workspace_bytes(ctypes.byref(cfg), 2)
    assert 0 < b1 < b2 and b1 % 256 == 0
    assert lib.te_vit_workspace_bytes(ctypes.byref(cfg), 0) < 0
    bad = vit_config(embed_dim=770)
    assert lib.te_vit_workspace_bytes(ctypes.byref(bad), 1) < 0
    assert b"invalid" in lib.te_last_error()
    assert lib.te_rollout_workspace_bytes(12, 2, 197) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the loud failure on a GPU-less host")
def test_no_cpu_fallback():
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import VisionTransformer
    from transformer_explainability_b200 import ops
    m = VisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=1, num_heads=4, num_classes=10)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(ValueError):
        ops.linear_relprop(torch.zeros(2, 4), torch.zeros(3, 4), torch.zeros(2, 3))


def test_reference_import_paths_resolve():
    import transformer_explainability_b200 as te
    te.install_aliases()
    from baselines.ViT.ViT_LRP import vit_base_patch16_224, compute_rollout_attention   # noqa: F401
    from baselines.ViT.ViT_explanation_generator import LRP                             # noqa: F401
    from modules.layers_ours import Linear, Add, Clone, einsum, IndexSelect, safe_divide  # noqa: F401
    import inspect
    sig = inspect.signature(LRP.generate_LRP)
    assert list(sig.parameters) == ["self", "input", "index", "method", "is_ablation", "start_layer"]
