"""results.hdf5 of generate_visualizations.py:27-102: the built-in minimal HDF5 emitter against the independent
spec-following reader (no HDF5 library exists in this image — see transformer_explainability_b200/hdf5_writer.py), and
the batched driver on the GPU."""
import os
import struct

import numpy as np
import pytest
import torch

from transformer_explainability_b200 import hdf5_writer as hw


def test_minimal_hdf5_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    w = hw.ResultsWriter(str(tmp_path), size=32, backend="builtin")
    imgs, viss, tgts = [], [], []
    for b in (1, 3, 2):                                                         # ragged batches, like a data loader's tail
        img = rng.standard_normal((b, 3, 32, 32)).astype(np.float32)
        vis = rng.random((b, 1, 32, 32)).astype(np.float32)
        tgt = rng.integers(0, 1000, size=b).astype(np.int64)
        w.append(img, vis, tgt)
        imgs.append(img); viss.append(vis); tgts.append(tgt)
    path = w.close()
    assert os.path.basename(path) == "results.hdf5"
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)                      # end-of-file address
    assert not [f for f in os.listdir(tmp_path) if f.startswith("te_h5_")]       # temporaries removed
    got = hw.read_minimal_hdf5(path)
    assert sorted(got) == ["image", "target", "vis"]                            # the names dataset/expl_hdf5.py:23-28 reads
    assert got["image"].dtype == np.float32 and got["vis"].dtype == np.float32 and got["target"].dtype == np.int32
    assert got["image"].shape == (6, 3, 32, 32) and got["vis"].shape == (6, 1, 32, 32) and got["target"].shape == (6,)
    assert np.array_equal(got["image"], np.concatenate(imgs))
    assert np.array_equal(got["vis"], np.concatenate(viss))
    assert np.array_equal(got["target"], np.concatenate(tgts).astype(np.int32))


def test_minimal_hdf5_rejects_bad_input(tmp_path):
    w = hw.ResultsWriter(str(tmp_path), size=16, backend="builtin")
    with pytest.raises(ValueError):
        w.append(np.zeros((2, 3, 16, 16)), np.zeros((2, 1, 8, 8)), np.zeros(2))
    w.append(np.zeros((1, 3, 16, 16)), np.ones((1, 1, 16, 16)), np.array([7]))
    got = hw.read_minimal_hdf5(w.close())
    assert got["target"].tolist() == [7] and float(got["vis"].min()) == 1.0
    with pytest.raises(TypeError):
        hw.write_minimal_hdf5(str(tmp_path / "x.h5"), {"a": ((2,), np.float64, np.zeros(2))})


def test_normalize_matches_reference_formula():
    x = torch.rand(2, 3, 4, 4)
    assert torch.allclose(hw.normalize(x), (x - 0.5) / 0.5)


@pytest.mark.gpu
def test_compute_saliency_and_save_on_the_engine(tmp_path):
    """The batched loop writes what per-sample calls produce: vis = per-sample min-max of the x16 bilinear map."""
    from oracle import vit as ovit
    from transformer_explainability_b200.baselines.ViT.ViT_LRP import vit_base_patch16_224
    from transformer_explainability_b200.baselines.ViT.ViT_explanation_generator import LRP
    params, heads = ovit.init_params("vit_base_patch16_224", seed=0)
    model = vit_base_patch16_224()
    model.load_state_dict(params)
    model = model.cuda().eval()
    lrp = LRP(model)
    g = torch.Generator().manual_seed(0)
    loader = [(torch.rand(2, 3, 224, 224, generator=g), torch.tensor([1, 2])),
              (torch.rand(1, 3, 224, 224, generator=g), torch.tensor([3]))]
    path = hw.compute_saliency_and_save(loader, str(tmp_path), "transformer_attribution", lrp=lrp, backend="builtin")
    got = hw.read_minimal_hdf5(path)
    assert got["vis"].shape == (3, 1, 224, 224) and got["image"].shape == (3, 3, 224, 224)
    assert got["target"].tolist() == [1, 2, 3]
    assert np.array_equal(got["image"][:2], loader[0][0].numpy())
    for s, (data, _) in ((0, loader[0]), (2, loader[1])):
        one = lrp.generate_LRP(hw.normalize(data[:1].cuda()), start_layer=1).reshape(1, 1, 14, 14)
        up = torch.nn.functional.interpolate(one, scale_factor=16, mode="bilinear")
        up = (up - up.min()) / (up.max() - up.min())
        assert np.abs(got["vis"][s] - up[0].cpu().numpy()).max() < 1e-4
    assert float(got["vis"].min()) >= 0.0 and float(got["vis"].max()) <= 1.0 + 1e-6
