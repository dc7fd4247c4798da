"""generate_visualization (example.ipynb:55-66) pieces: 14x14 -> bilinear x16 -> per-sample min-max."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from transformer_explainability_b200.visualization import relevance_to_heatmap


def _notebook_reference(row):
    """The notebook's own sequence for one explanation (example.ipynb:57-60)."""
    t = row.reshape(1, 1, 14, 14)
    t = F.interpolate(t, scale_factor=16, mode='bilinear')
    t = t.reshape(224, 224)
    return (t - t.min()) / (t.max() - t.min())


def test_heatmap_has_no_host_path():
    with pytest.raises(ValueError):
        relevance_to_heatmap(torch.rand(2, 196))


@pytest.mark.gpu
def test_heatmap_kernel_matches_notebook_sequence():
    g = torch.Generator().manual_seed(3)
    maps = torch.rand(7, 196, generator=g) * 3e-4
    heat = relevance_to_heatmap(maps.cuda())
    assert heat.shape == (7, 224, 224) and heat.is_cuda
    for s in range(7):
        assert (heat[s].cpu() - _notebook_reference(maps[s])).abs().max() < 2e-6
    assert float(heat.min()) == 0.0 and float(heat.max()) == 1.0


@pytest.mark.gpu
def test_generate_visualization_end_to_end_gpu():
    cv2 = pytest.importorskip("cv2")
    from oracle import vit as ovit
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import VisionTransformer
    from transformer_explainability_b200.baselines.ViT.ViT_explanation_generator import LRP
    from transformer_explainability_b200.visualization import generate_visualization
    params, heads = ovit.init_params("vit_base_patch16_224", seed=0)
    m = VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True)
    m.load_state_dict(params)
    m = m.cuda().eval()
    img = torch.rand(3, 224, 224, generator=torch.Generator().manual_seed(1))
    vis = generate_visualization(LRP(m), img)
    assert vis.shape == (224, 224, 3) and vis.dtype == np.uint8


def test_eraser_rationale_lines_follow_reference_format():
    """bert_pipeline.py:563-582: top-k token rationales, one JSON line per document and k, cumulative list."""
    import json
    import torch
    from transformer_explainability_b200 import eraser
    scores = torch.tensor([[0.1, -3.0, 0.9, 0.5, 0.2, 0.0, 0.7, 0.3]])
    lines = eraser.rationale_lines(["doc_a"], scores, ks=(2, 3))
    first = json.loads(lines[2][0])
    assert first["annotation_id"] == "doc_a" and first["rationales"][0]["docid"] == "doc_a"
    assert first["rationales"][0]["hard_rationale_predictions"] == [{"start_token": 2, "end_token": 3},
                                                                    {"start_token": 6, "end_token": 7}]
    second = json.loads(lines[3][0])["rationales"][0]["hard_rationale_predictions"]
    assert [r["start_token"] for r in second] == [2, 6, 2, 6, 3]          # the reference never resets the list
