"""Drop-in for ``BERT_explainability/modules/BERT/ExplanationGenerator.py`` (class ``Generator``, ``generate_LRP``).

The other generators of the reference file (partial/full LRP, raw rollout, attention GradCAM ...) are comparison
baselines and out of scope."""
import torch

from transformer_explainability_b200 import ops


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

    def generate_LRP_batched(self, input_ids, attention_mask=None, index=None, start_layer=11, chunk=None,
                             return_index=False):
        """B independent sequences of equal length in one engine call: [B,S] -> [B,S]."""
        maps, idx = self.model.engine().explain(input_ids, attention_mask, index=index, start_layer=start_layer,
                                                chunk=chunk)
        return (maps, idx) if return_index else maps
