"""Drop-in for ``BERT_explainability/modules/BERT/BertForSequenceClassification.py``.

``BertForSequenceClassification(config)`` is a parameter container with the HF ``state_dict`` keys
(``bert.embeddings.*``, ``bert.encoder.layer.{i}.attention.self.{query,key,value}.*`` ...,
``bert.pooler.dense.*``, ``classifier.*``); ``forward`` and ``relprop`` run on the CUDA engine
(``engine.BertEngine``) for a batch of independent sequences.  ``config`` is a ``transformers.BertConfig`` or any
object with the same attribute names.  No ``param.grad`` side effect, no autograd graph.
"""
import torch
import torch.nn as nn

from transformer_explainability_b200 import _lib
from transformer_explainability_b200.engine import BertEngine, bert_config


class _SelfAttention(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.query = nn.Linear(d, d)
        self.key = nn.Linear(d, d)
        self.value = nn.Linear(d, d)
        self._owner = None
        self._layer = -1

    # accessors of BERT.py:281-297, served from the engine workspace
    def get_attn(self):
        return self._owner[0]._engine_tensor("attn", self._layer)

    def get_attn_cam(self):
        return self._owner[0]._engine_tensor("attn_cam", self._layer)

    def get_attn_gradients(self):
        return self._owner[0]._engine_tensor("attn_grad", self._layer)


class _DenseLN(nn.Module):
    def __init__(self, i, o, eps):
        super().__init__()
        self.dense = nn.Linear(i, o)
        self.LayerNorm = nn.LayerNorm(o, eps=eps)


class _Dense(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class _Attention(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.self = _SelfAttention(d)
        self.output = _DenseLN(d, d, eps)


class _Layer(nn.Module):
    def __init__(self, d, f, eps):
        super().__init__()
        self.attention = _Attention(d, eps)
        self.intermediate = _Dense(d, f)
        self.output = _DenseLN(f, d, eps)


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(cfg.max_position_embeddings).expand((1, -1)))


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(cfg.hidden_size, cfg.intermediate_size, cfg.layer_norm_eps)
                                    for _ in range(cfg.num_hidden_layers)])


class _BertModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)
        self.pooler = _Dense(cfg.hidden_size, cfg.hidden_size)


class BertForSequenceClassification(nn.Module):
    def __init__(self, config):
        super().__init__()
        if getattr(config, "hidden_act", "gelu") != "gelu":
            raise NotImplementedError("only hidden_act='gelu' (the BertConfig default) is implemented")
        self.config = config
        self.num_labels = config.num_labels
        self.bert = _BertModel(config)
        self.classifier = nn.Linear(config.hidden_size, config.num_labels)
        for i, l in enumerate(self.bert.encoder.layer):
            l.attention.self._owner = (self,)
            l.attention.self._layer = i
        self._cfg = bert_config(config.vocab_size, config.max_position_embeddings, config.type_vocab_size,
                                config.hidden_size, config.num_hidden_layers, config.num_attention_heads,
                                config.intermediate_size, config.num_labels, config.layer_norm_eps)
        self._engine = None
        self._weights_version = None
        self.engine_flags = 0
        std = getattr(config, "initializer_range", 0.02)
        for m in self.modules():                                   # transformers' _init_weights
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, mean=0.0, std=std)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)

    def _version(self):
        return tuple(p._version for p in self.parameters()) + (str(self.classifier.weight.device),)

    def engine(self):
        dev = self.classifier.weight.device
        if dev.type != "cuda":
            raise RuntimeError("the B200 engine has no CPU path: move the model to a CUDA device (model.cuda())")
        v = self._version()
        if self._engine is None or self._engine.device != dev:
            self._engine = BertEngine(self._cfg, device=dev, flags=self.engine_flags)
            self._weights_version = None
        if self._weights_version != v:
            self._engine.load_state_dict(self.state_dict())
            self._weights_version = v
        self._engine.flags = self.engine_flags
        return self._engine

    def _engine_tensor(self, name, layer):
        if self._engine is None or self._engine.last[0] <= 0:
            raise RuntimeError("no saved activations: call model(input_ids, attention_mask) first")
        return self._engine.tensor(name, layer)

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        """``BertForSequenceClassification.forward`` (:23-81) with return_dict=False: returns ``(logits,)``."""
        if token_type_ids is not None or position_ids is not None or head_mask is not None or inputs_embeds is not None:
            raise NotImplementedError("only input_ids / attention_mask are used on the attribution path "
                                      "(bert_pipeline.py:443,551)")
        return (self.engine().forward(input_ids, attention_mask),)

    def relprop(self, cam=None, **kwargs):
        """``relprop`` (:83-88): relevance at the encoder input [B,S,D]; leaves attn_cam / attn_gradients of every
        layer readable through ``layer.attention.self.get_attn_cam()`` ... like the reference."""
        if kwargs.get("alpha", 1) != 1:
            raise NotImplementedError("only alpha=1 is implemented (the only value the reference passes)")
        eng = self.engine()
        index = cam.argmax(dim=-1).to(torch.int32) if cam is not None else None
        eng.attribute(index=index, start_layer=0, flags=eng.flags | _lib.FLAG_RELPROP_TO_INPUT)
        return eng.tensor("relevance_in")
