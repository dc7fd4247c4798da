"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): tiny ViT, tiny BERT, and every tensor-core
kernel at its smallest supported shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vit as ovit, bert as obert                       # noqa: E402
from transformer_explainability_b200 import _lib, ops               # noqa: E402
from transformer_explainability_b200.baselines.ViT.ViT_LRP import VisionTransformer   # noqa: E402


def main():
    params, heads = ovit.init_params("vit_tiny_test", seed=1, rand_affine=True)
    m = VisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=3, num_heads=heads, mlp_ratio=4., qkv_bias=True,
                          num_classes=10)
    m.load_state_dict(params)
    m = m.cuda().eval()
    x = torch.randn(3, 3, 32, 32, generator=torch.Generator().manual_seed(5)).cuda()
    for fl in (0, _lib.FLAG_ROLLOUT_FUSED):
        maps, idx = m.engine().explain(x, flags=fl)
    torch.cuda.synchronize()
    print("tiny vit ok", float(maps.abs().max()))
    # a ViT with tensor-core-eligible dims (dim 256, heads 4 -> head_dim 64, mlp 256): all fast paths, 2 blocks
    big = VisionTransformer(img_size=32, patch_size=8, embed_dim=256, depth=2, num_heads=4, mlp_ratio=1., qkv_bias=True,
                            num_classes=12).cuda().eval()
    maps, idx = big.engine().explain(x, flags=_lib.FLAG_ALL_FAST)
    torch.cuda.synchronize()
    print("tc vit ok", float(maps.abs().max()))
    from transformers import BertConfig
    from transformer_explainability_b200.BERT_explainability.modules.BERT.BertForSequenceClassification import \
        BertForSequenceClassification
    bp, bh = obert.init_params(seed=3, vocab=100, max_pos=32, dim=64, depth=2, heads=4, inter=128, rand_affine=True)
    bm = BertForSequenceClassification(BertConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                                                  intermediate_size=128, vocab_size=100, max_position_embeddings=32,
                                                  num_labels=2))
    bm.load_state_dict(bp, strict=False)
    bm = bm.cuda().eval()
    ids = torch.randint(5, 100, (2, 19), generator=torch.Generator().manual_seed(1)).cuda()
    mask = torch.ones_like(ids)
    mask[1, 12:] = 0
    maps, idx = bm.engine().explain(ids, mask, start_layer=0, flags=_lib.FLAG_ROLLOUT_FUSED)
    torch.cuda.synchronize()
    print("tiny bert ok", float(maps.abs().max()))
    g = torch.Generator(device="cuda").manual_seed(0)
    xx = torch.randn(130, 256, device="cuda", generator=g)
    ww = torch.randn(256, 256, device="cuda", generator=g) * 0.05
    rr = torch.rand(130, 256, device="cuda", generator=g)
    ops.linear_relprop(xx, ww, rr, tensor_cores=True)
    ops.linear_forward(xx, ww, None, tensor_cores=True)
    ops.linear_backward(rr, ww, tensor_cores=True)
    torch.cuda.synchronize()
    print("tc kernels ok")


if __name__ == "__main__":
    main()
