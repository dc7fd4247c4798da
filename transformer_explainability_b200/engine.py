"""Host-side engine objects: flat frozen-weight buffer, workspace, and the explain call.

The engine is the B200 replacement for the stateful hook machinery of the reference
(``modules/layers_ours.py:16-27`` forward hooks, ``retain_graph=True`` autograd graph): it owns one
flat fp32 weight buffer (the unit of the single NCCL broadcast) and one activation workspace per
stream, and issues O(1) launches per block per BATCH through the C ABI.
"""
import ctypes
import functools

import torch

from . import _lib
from ._lib import TeVitConfig, check, ptr


def _on_engine_device(fn):
    """Make the engine's device current for the duration of the call: the C library launches on the current device
    and neither it nor ``torch.cuda.current_stream(dev)`` switches devices (a model moved to ``cuda:1`` without
    ``torch.cuda.set_device(1)`` would otherwise fail with an invalid resource handle)."""
    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapped


def vit_config(img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
               mlp_ratio=4., distilled=False, eps_block=1e-6, eps_final=1e-5):
    return TeVitConfig(img_size, patch_size, in_chans, num_classes, embed_dim, depth, num_heads,
                       int(embed_dim * mlp_ratio), int(bool(distilled)), eps_block, eps_final)


class ViTEngine:
    """Runs ``generate_LRP(method='transformer_attribution')`` for batches of independent inputs."""

    def __init__(self, cfg, state_dict=None, device=None, flags=0):
        if not torch.cuda.is_available():
            raise RuntimeError("transformer_explainability_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.flags = flags
        n = check(self.lib.te_vit_num_weights(ctypes.byref(cfg)), "te_vit_num_weights")
        self.weight_table = []
        for i in range(n):
            self.weight_table.append((self.lib.te_vit_weight_name(ctypes.byref(cfg), i).decode(),
                                      self.lib.te_vit_weight_numel(ctypes.byref(cfg), i),
                                      self.lib.te_vit_weight_offset(ctypes.byref(cfg), i)))
        total = check(self.lib.te_vit_weight_total(ctypes.byref(cfg)), "te_vit_weight_total")
        self.weights = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.derived = None                      # tensor-core weight copies, built on demand
        self._ws = None
        self._ws_batch = 0
        self.tokens = (cfg.img_size // cfg.patch_size) ** 2 + (2 if cfg.distilled else 1)
        self.prefix = 2 if cfg.distilled else 1
        self.last_batch = 0
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ---- weights ------------------------------------------------------------------------------
    @_on_engine_device
    def load_state_dict(self, sd):
        """Pack a reference-keyed ``state_dict`` (timm ViT names) into the flat device buffer."""
        host = torch.zeros(self.weights.numel(), dtype=torch.float32)
        for name, numel, off in self.weight_table:
            if name not in sd:
                if name.endswith("qkv.bias"):
                    continue                          # qkv_bias=False models: zeros
                raise KeyError("state_dict is missing %r" % name)
            t = sd[name].detach().to(torch.float32).reshape(-1).cpu()
            if t.numel() != numel:
                raise ValueError("%s: expected %d values, got %d" % (name, numel, t.numel()))
            host[off:off + numel] = t
        self.weights.copy_(host, non_blocking=False)
        self.derived = None

    @_on_engine_device
    def _derived(self, flags):
        """W+/W-/W+^T/W-^T TF32 copies for the tcgen05 z+ path (built once per weight load)."""
        if not (flags & _lib.FLAG_TENSOR_CORES):
            return None
        if self.derived is None:
            n = check(self.lib.te_vit_derived_total(ctypes.byref(self.cfg)), "te_vit_derived_total")
            self.derived = torch.empty(n, dtype=torch.float32, device=self.device)
            check(self.lib.te_vit_prepare_derived(ctypes.byref(self.cfg), ptr(self.weights), ptr(self.derived),
                                                  self._stream()), "te_vit_prepare_derived")
        return self.derived

    def broadcast_weights(self, src=0, group=None):
        """The one collective of the path: NCCL broadcast of the flat frozen-weight buffer."""
        import torch.distributed as dist
        dist.broadcast(self.weights, src=src, group=group)

    # ---- workspace ----------------------------------------------------------------------------
    def workspace_bytes(self, batch):
        return check(self.lib.te_vit_workspace_bytes(ctypes.byref(self.cfg), batch), "te_vit_workspace_bytes")

    def _workspace(self, batch):
        if self._ws is None or self._ws_batch != batch:
            self._ws = None
            nbytes = self.workspace_bytes(batch)
            self._ws = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            self._ws_batch = batch
        return self._ws

    def max_chunk(self, limit=None, reserve_bytes=4 << 30):
        """Largest per-call batch whose workspace fits in free HBM (activations of all blocks are kept)."""
        free, _ = torch.cuda.mem_get_info(self.device)
        if self._ws is not None:
            free += self._ws.numel() * 4
        per = self.workspace_bytes(2) - self.workspace_bytes(1)
        b = max(1, int((free - reserve_bytes) // max(per, 1)))
        return min(b, limit) if limit else b

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- the three calls ------------------------------------------------------------------------
    @_on_engine_device
    def forward(self, images, flags=None):
        """``model(x)``: logits [B,C]; leaves the activations in the workspace."""
        images = images.to(self.device, torch.float32).contiguous()
        b = images.shape[0]
        ws = self._workspace(b)
        fl = self.flags if flags is None else flags
        logits = torch.empty(b, self.cfg.num_classes, dtype=torch.float32, device=self.device)
        check(self.lib.te_vit_forward(ctypes.byref(self.cfg), ptr(self.weights), ptr(self._derived(fl)), ptr(images), b,
                                      fl, ptr(logits), ptr(ws), ws.numel() * 4, self._stream()), "te_vit_forward")
        self.last_batch = b
        self._last_images = images
        return logits

    @_on_engine_device
    def relprop_pixels(self, index=None, per_channel=False, flags=None):
        """``method="full"`` (ViT_LRP.py:337-343) on the activations of the last ``forward``: the relprop is run to
        the encoder input, through ``self.add`` and the patch convolution's z^B rule.  Returns the relevance of every
        pixel, [B,H,W] (channels summed, what the reference returns) or [B,C,H,W] with ``per_channel``."""
        fl = (self.flags if flags is None else flags) | _lib.FLAG_RELPROP_TO_INPUT
        self.attribute(index=index, start_layer=0, flags=fl)
        b = self.last_batch
        images = getattr(self, "_last_images", None)
        if images is None or images.shape[0] != b:
            raise RuntimeError("relprop_pixels() needs the images of the preceding forward()")
        ws = self._workspace(b)
        c, s = self.cfg.in_chans, self.cfg.img_size
        out = torch.empty((b, c, s, s) if per_channel else (b, s, s), dtype=torch.float32, device=self.device)
        check(self.lib.te_vit_relprop_pixels_ex(ctypes.byref(self.cfg), ptr(self.weights), ptr(images), b, fl,
                                                None if per_channel else ptr(out), ptr(out) if per_channel else None,
                                                ptr(ws), ws.numel() * 4, self._stream()), "te_vit_relprop_pixels_ex")
        return out

    @_on_engine_device
    def attribute(self, index=None, start_layer=0, flags=None):
        """Backward + relprop + rollout on the activations of the last ``forward``.
        Returns (maps [B,N-prefix], index [B] int32)."""
        b = self.last_batch
        if b <= 0:
            raise RuntimeError("attribute() needs a preceding forward()")
        ws = self._workspace(b)
        idx = self._index_tensor(index, b)
        maps = torch.empty(b, self.tokens - self.prefix, dtype=torch.float32, device=self.device)
        fl = self.flags if flags is None else flags
        check(self.lib.te_vit_attribute(ctypes.byref(self.cfg), ptr(self.weights), ptr(self._derived(fl)), b, ptr(idx),
                                        int(start_layer), fl, ptr(maps), ptr(ws), ws.numel() * 4, self._stream()),
              "te_vit_attribute")
        return maps, idx

    @_on_engine_device
    def explain(self, images, index=None, start_layer=0, flags=None, chunk=None, return_logits=False):
        """``LRP.generate_LRP`` for a batch of independent inputs (device-resident in, device-resident out)."""
        images = images.to(self.device, torch.float32).contiguous()
        B = images.shape[0]
        chunk = min(B, chunk or self.max_chunk(limit=B))
        maps = torch.empty(B, self.tokens - self.prefix, dtype=torch.float32, device=self.device)
        idx_all = self._index_tensor(index, B)
        logits = torch.empty(B, self.cfg.num_classes, dtype=torch.float32, device=self.device) if return_logits else None
        fl = self.flags if flags is None else flags
        derived = self._derived(fl)
        for s in range(0, B, chunk):
            e = min(B, s + chunk)
            ws = self._workspace(chunk if e - s == chunk else e - s)
            check(self.lib.te_vit_explain(ctypes.byref(self.cfg), ptr(self.weights), ptr(derived), ptr(images[s:e]), e - s,
                                          ptr(idx_all[s:e]), int(start_layer), fl, ptr(maps[s:e]),
                                          ptr(logits[s:e]) if logits is not None else None, ptr(ws), ws.numel() * 4,
                                          self._stream()), "te_vit_explain")
            self.last_batch = e - s
            self._last_images = images[s:e]
        if return_logits:
            return maps, idx_all, logits
        return maps, idx_all

    # ---- CUDA-graph replay of the fixed-shape step ------------------------------------------------
    @_on_engine_device
    def explain_graphed(self, images, index=None, start_layer=0, flags=None, return_logits=False):
        """``explain`` with the whole step (~40 launches per block) captured once in a CUDA graph and replayed: the
        shapes, the workspace and every kernel argument are fixed for a given (batch, start_layer, flags), so small
        per-GPU batches (a fixed global batch sharded over 8 GPUs) are not bound by launch gaps.
        Inputs are copied into static buffers; outputs are views of static buffers (overwritten by the next call)."""
        images = images.to(self.device, torch.float32)
        B = images.shape[0]
        fl = self.flags if flags is None else flags
        key = (B, int(start_layer), int(fl), tuple(images.shape[1:]))
        g = getattr(self, "_graphs", None)
        if g is None:
            g = self._graphs = {}
        if key not in g:
            if len(g) >= 4:
                g.clear()                                   # bounded cache: graphs pin their static buffers
            ws = self._workspace(B)
            derived = self._derived(fl)
            st = dict(images=torch.empty_like(images, memory_format=torch.contiguous_format),
                      idx_in=torch.full((B,), -1, dtype=torch.int32, device=self.device),
                      idx=torch.full((B,), -1, dtype=torch.int32, device=self.device),
                      maps=torch.empty(B, self.tokens - self.prefix, dtype=torch.float32, device=self.device),
                      logits=torch.empty(B, self.cfg.num_classes, dtype=torch.float32, device=self.device), ws=ws)
            st["images"].copy_(images)

            def run():
                st["idx"].copy_(st["idx_in"])
                check(self.lib.te_vit_explain(ctypes.byref(self.cfg), ptr(self.weights), ptr(derived), ptr(st["images"]), B,
                                              ptr(st["idx"]), int(start_layer), fl, ptr(st["maps"]), ptr(st["logits"]),
                                              ptr(ws), ws.numel() * 4, self._stream()), "te_vit_explain")

            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                run()                                       # warm-up outside capture: per-device kernel attributes, allocator
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            st["graph"] = graph
            g[key] = st
        st = g[key]
        if st["ws"] is not self._ws:                         # the workspace was re-allocated for another batch size
            del g[key]
            return self.explain_graphed(images, index=index, start_layer=start_layer, flags=flags,
                                        return_logits=return_logits)
        st["images"].copy_(images, non_blocking=True)
        st["idx_in"].copy_(self._index_tensor(index, B))
        st["graph"].replay()
        self.last_batch = B
        self._last_images = st["images"]
        if return_logits:
            return st["maps"], st["idx"], st["logits"]
        return st["maps"], st["idx"]

    def _index_tensor(self, index, b):
        if index is None:
            return torch.full((b,), -1, dtype=torch.int32, device=self.device)
        t = torch.as_tensor(index, device=self.device).reshape(-1).to(torch.int32)
        if t.numel() == 1 and b > 1:
            t = t.expand(b)
        if t.numel() != b:
            raise ValueError("index must have one entry per sample")
        return t.contiguous().clone()

    # ---- accessors (get_attn / get_attn_gradients / get_attn_cam ..., ViT_LRP.py:102-130) ---------
    def tensor(self, name, layer=0):
        b = self.last_batch
        ws = self._workspace(b)
        p = ctypes.c_void_p()
        dims = (ctypes.c_longlong * 4)()
        strides = (ctypes.c_longlong * 4)()
        check(self.lib.te_vit_tensor(ctypes.byref(self.cfg), b, ptr(ws), name.encode(), layer, ctypes.byref(p), dims,
                                     strides), "te_vit_tensor")
        off = (p.value - ws.data_ptr()) // 4
        nd = 4
        while nd > 2 and dims[nd - 1] == 1:
            nd -= 1
        return torch.as_strided(ws, [int(dims[i]) for i in range(nd)], [int(strides[i]) for i in range(nd)], off)


def bert_config(vocab_size=30522, max_position_embeddings=512, type_vocab_size=2, hidden_size=768,
                num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, num_labels=2,
                layer_norm_eps=1e-12):
    from ._lib import TeBertConfig
    return TeBertConfig(vocab_size, max_position_embeddings, type_vocab_size, hidden_size, num_hidden_layers,
                        num_attention_heads, intermediate_size, num_labels, layer_norm_eps)


class BertEngine:
    """``Generator.generate_LRP`` (BERT_explainability/modules/BERT/ExplanationGenerator.py:28-59) for batches of
    independent sequences of one length."""

    def __init__(self, cfg, state_dict=None, device=None, flags=0):
        if not torch.cuda.is_available():
            raise RuntimeError("transformer_explainability_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.flags = flags
        n = check(self.lib.te_bert_num_weights(ctypes.byref(cfg)), "te_bert_num_weights")
        self.weight_table = [(self.lib.te_bert_weight_name(ctypes.byref(cfg), i).decode(),
                              self.lib.te_bert_weight_numel(ctypes.byref(cfg), i),
                              self.lib.te_bert_weight_offset(ctypes.byref(cfg), i)) for i in range(n)]
        total = check(self.lib.te_bert_weight_total(ctypes.byref(cfg)), "te_bert_weight_total")
        self.weights = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.derived = None
        self._ws = None
        self._ws_key = None
        self.last = (0, 0)
        if state_dict is not None:
            self.load_state_dict(state_dict)

    @_on_engine_device
    def load_state_dict(self, sd):
        host = torch.zeros(self.weights.numel(), dtype=torch.float32)
        for name, numel, off in self.weight_table:
            if name not in sd:
                raise KeyError("state_dict is missing %r" % name)
            t = sd[name].detach().to(torch.float32).reshape(-1).cpu()
            if t.numel() != numel:
                raise ValueError("%s: expected %d values, got %d" % (name, numel, t.numel()))
            host[off:off + numel] = t
        self.weights.copy_(host)
        self.derived = None

    def broadcast_weights(self, src=0, group=None):
        import torch.distributed as dist
        dist.broadcast(self.weights, src=src, group=group)

    @_on_engine_device
    def _derived(self, flags):
        if not (flags & _lib.FLAG_TENSOR_CORES):
            return None
        if self.derived is None:
            n = check(self.lib.te_bert_derived_total(ctypes.byref(self.cfg)), "te_bert_derived_total")
            self.derived = torch.empty(n, dtype=torch.float32, device=self.device)
            check(self.lib.te_bert_prepare_derived(ctypes.byref(self.cfg), ptr(self.weights), ptr(self.derived),
                                                   self._stream()), "te_bert_prepare_derived")
        return self.derived

    def workspace_bytes(self, batch, seq):
        return check(self.lib.te_bert_workspace_bytes(ctypes.byref(self.cfg), batch, seq), "te_bert_workspace_bytes")

    def _workspace(self, batch, seq):
        if self._ws is None or self._ws_key != (batch, seq):
            self._ws = None
            self._ws = torch.empty(self.workspace_bytes(batch, seq) // 4, dtype=torch.float32, device=self.device)
            self._ws_key = (batch, seq)
        return self._ws

    def max_chunk(self, seq, limit=None, reserve_bytes=4 << 30):
        free, _ = torch.cuda.mem_get_info(self.device)
        if self._ws is not None:
            free += self._ws.numel() * 4
        per = self.workspace_bytes(2, seq) - self.workspace_bytes(1, seq)
        b = max(1, int((free - reserve_bytes) // max(per, 1)))
        return min(b, limit) if limit else b

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ids(self, input_ids, attention_mask):
        if not input_ids.is_cuda and input_ids.numel() and (int(input_ids.min()) < 0 or
                                                            int(input_ids.max()) >= self.cfg.vocab_size):
            raise ValueError("input_ids outside [0, vocab_size)")     # device-resident ids: the kernel writes NaN rows
        ids = input_ids.to(self.device, torch.int64).contiguous()
        if attention_mask is None:
            attention_mask = torch.ones_like(ids)
        return ids, attention_mask.to(self.device, torch.int64).contiguous()

    @_on_engine_device
    def forward(self, input_ids, attention_mask=None, flags=None):
        ids, mask = self._ids(input_ids, attention_mask)
        b, s = ids.shape
        ws = self._workspace(b, s)
        fl = self.flags if flags is None else flags
        logits = torch.empty(b, self.cfg.num_labels, dtype=torch.float32, device=self.device)
        check(self.lib.te_bert_forward(ctypes.byref(self.cfg), ptr(self.weights), ptr(self._derived(fl)), ptr(ids),
                                       ptr(mask), b, s, fl, ptr(logits), ptr(ws), ws.numel() * 4, self._stream()),
              "te_bert_forward")
        self.last = (b, s)
        return logits

    @_on_engine_device
    def attribute(self, index=None, start_layer=11, flags=None):
        b, s = self.last
        if b <= 0:
            raise RuntimeError("attribute() needs a preceding forward()")
        ws = self._workspace(b, s)
        idx = ViTEngine._index_tensor(self, index, b)
        maps = torch.empty(b, s, dtype=torch.float32, device=self.device)
        fl = self.flags if flags is None else flags
        check(self.lib.te_bert_attribute(ctypes.byref(self.cfg), ptr(self.weights), ptr(self._derived(fl)), b, s, ptr(idx),
                                         int(start_layer), fl, ptr(maps), ptr(ws), ws.numel() * 4, self._stream()),
              "te_bert_attribute")
        return maps, idx

    @_on_engine_device
    def explain(self, input_ids, attention_mask=None, index=None, start_layer=11, flags=None, chunk=None,
                return_logits=False):
        ids, mask = self._ids(input_ids, attention_mask)
        B, S = ids.shape
        chunk = min(B, chunk or self.max_chunk(S, limit=B))
        maps = torch.empty(B, S, dtype=torch.float32, device=self.device)
        idx_all = ViTEngine._index_tensor(self, index, B)
        logits = torch.empty(B, self.cfg.num_labels, dtype=torch.float32, device=self.device) if return_logits else None
        fl = self.flags if flags is None else flags
        derived = self._derived(fl)
        for s0 in range(0, B, chunk):
            e = min(B, s0 + chunk)
            ws = self._workspace(e - s0, S)
            check(self.lib.te_bert_explain(ctypes.byref(self.cfg), ptr(self.weights), ptr(derived), ptr(ids[s0:e]),
                                           ptr(mask[s0:e]), e - s0, S, ptr(idx_all[s0:e]), int(start_layer), fl,
                                           ptr(maps[s0:e]), ptr(logits[s0:e]) if logits is not None else None, ptr(ws),
                                           ws.numel() * 4, self._stream()), "te_bert_explain")
            self.last = (e - s0, S)
        if return_logits:
            return maps, idx_all, logits
        return maps, idx_all

    def tensor(self, name, layer=0):
        b, s = self.last
        ws = self._workspace(b, s)
        p = ctypes.c_void_p()
        dims = (ctypes.c_longlong * 4)()
        strides = (ctypes.c_longlong * 4)()
        check(self.lib.te_bert_tensor(ctypes.byref(self.cfg), b, s, ptr(ws), name.encode(), layer, ctypes.byref(p), dims,
                                      strides), "te_bert_tensor")
        off = (p.value - ws.data_ptr()) // 4
        nd = 4
        while nd > 2 and dims[nd - 1] == 1:
            nd -= 1
        return torch.as_strided(ws, [int(dims[i]) for i in range(nd)], [int(strides[i]) for i in range(nd)], off)
