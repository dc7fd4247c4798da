"""Drop-in for ``BERT_explainability/modules/BERT/ExplanationGenerator.py`` (class ``Generator``).

``generate_LRP`` is the transformer-attribution hot path; the comparison generators of the same class
(``generate_LRP_last_layer``, ``generate_full_lrp``, ``generate_attn_last_layer``, ``generate_rollout``,
``generate_attn_gradcam``, reference ``:61-155``) are served from the same engine passes and the same kernels.
Every generator accepts a batch of independent sequences of one length ([B,S] -> [B,S]); B = 1 is the reference call."""
import torch

from transformer_explainability_b200 import _lib, ops


def compute_rollout_attention(all_layer_matrices, start_layer=0):
    """``ExplanationGenerator.py:7-18``: (M + I) / rowsum, chained from ``start_layer``."""
    return ops.compute_rollout_attention(all_layer_matrices, start_layer=start_layer, normalize=True)


class Generator:
    def __init__(self, model):
        self.model = model
        self.model.eval()

    def forward(self, input_ids, attention_mask):
        return self.model(input_ids, attention_mask)

    def generate_LRP(self, input_ids, attention_mask, index=None, start_layer=11):
        """``:28-59``: [1,S] ids -> [1,S] token relevance (row 0 of the normalised rollout, element 0 = row minimum)."""
        maps, _ = self.model.engine().explain(input_ids, attention_mask, index=index, start_layer=start_layer)
        return maps

    # ---- comparison generators (reference :61-155) ---------------------------------------------------------------
    def _last(self):
        return self.model.bert.encoder.layer[-1].attention.self

    def _run(self, input_ids, attention_mask, index, start_layer, extra_flags):
        eng = self.model.engine()
        eng.forward(input_ids, attention_mask)
        eng.attribute(index=index, start_layer=start_layer, flags=eng.flags | extra_flags)
        return eng

    def generate_LRP_last_layer(self, input_ids, attention_mask, index=None):
        """``:61-83``: head-mean of the clamped attention relevance (attn_cam) of the last layer, row 0, [0] = 0."""
        eng = self._run(input_ids, attention_mask, index, self.model._cfg.layers - 1, 0)
        cam = ops.head_reduce(self._last().get_attn_cam(), mode="relu_mean")
        cam[:, 0, 0] = 0
        return cam[:, 0]

    def generate_full_lrp(self, input_ids, attention_mask, index=None):
        """``:85-105``: LRP carried to the encoder input, summed over the hidden dimension, [0] = 0."""
        eng = self._run(input_ids, attention_mask, index, 0, _lib.FLAG_RELPROP_TO_INPUT)
        cam = eng.tensor("relevance_in").sum(dim=2)
        cam[:, 0] = 0
        return cam

    def generate_attn_last_layer(self, input_ids, attention_mask, index=None):
        """``:107-113``: head-mean of the last layer's raw attention, row 0, [0] = 0."""
        self.model.engine().forward(input_ids, attention_mask)
        cam = ops.head_reduce(self._last().get_attn(), mode="mean")
        cam[:, 0, 0] = 0
        return cam[:, 0]

    def generate_rollout(self, input_ids, attention_mask, start_layer=0, index=None):
        """``:115-127``: rollout of the head-averaged raw attention, row 0, [0] = 0."""
        self.model.engine().forward(input_ids, attention_mask)
        mats = [ops.head_reduce(l.attention.self.get_attn(), mode="mean") for l in self.model.bert.encoder.layer]
        rollout = compute_rollout_attention(mats, start_layer=start_layer)
        rollout[:, 0, 0] = 0
        return rollout[:, 0]

    def generate_attn_gradcam(self, input_ids, attention_mask, index=None):
        """``:129-155``: last-layer attention weighted per head by its mean gradient, relu(mean over heads), min-max
        normalised over the [S,S] map, row 0, [0] = 0."""
        eng = self._run(input_ids, attention_mask, index, self.model._cfg.layers - 1, _lib.FLAG_GRADIENTS_ONLY)
        att = self._last()
        w = ops.head_region_mean(att.get_attn_gradients())
        cam = ops.head_reduce(att.get_attn(), head_weight=w, mode="mean_relu")
        lo = cam.amin(dim=(1, 2), keepdim=True)
        hi = cam.amax(dim=(1, 2), keepdim=True)
        cam = (cam - lo) / (hi - lo)
        cam[:, 0, 0] = 0
        return cam[:, 0]

    def generate_LRP_batched(self, input_ids, attention_mask=None, index=None, start_layer=11, chunk=None,
                             return_index=False):
        """B independent sequences of equal length in one engine call: [B,S] -> [B,S]."""
        maps, idx = self.model.engine().explain(input_ids, attention_mask, index=index, start_layer=start_layer,
                                                chunk=chunk)
        return (maps, idx) if return_index else maps
