"""Drop-in for ``baselines/ViT/ViT_explanation_generator.py`` (class ``LRP``).

``generate_LRP`` keeps the reference signature and return shape ([1, 196]); ``generate_LRP_batched`` is the
batched addition: B independent explanations per call, device-resident in and out.
The ``Baselines`` class (raw-attention rollout / attention GradCAM on the hook-free ViT) is out of scope.
"""
import torch


class LRP:
    def __init__(self, model):
        self.model = model
        self.model.eval()

    def generate_LRP(self, input, index=None, method="transformer_attribution", is_ablation=False, start_layer=0):
        """``ViT_explanation_generator.py:25-41``: forward, class pick, one-hot seed, relprop."""
        output = self.model(input)
        if index is None:
            index = output.argmax(dim=-1)
        index = torch.as_tensor(index, device=output.device).reshape(-1).long()
        one_hot = torch.zeros_like(output)
        one_hot[torch.arange(output.shape[0], device=output.device), index] = 1
        return self.model.relprop(one_hot, method=method, is_ablation=is_ablation, start_layer=start_layer, alpha=1)

    def generate_LRP_batched(self, input, index=None, start_layer=0, chunk=None, return_index=False):
        """B independent ``transformer_attribution`` explanations in one engine call: [B,3,H,W] -> [B,196]."""
        maps, idx = self.model.engine().explain(input, index=index, start_layer=start_layer, chunk=chunk)
        return (maps, idx) if return_index else maps
