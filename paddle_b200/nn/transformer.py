"""Transformer layers. Parity: python/paddle/nn/layer/transformer.py."""
from __future__ import annotations

import collections
import copy

import torch

from ..tensor import Tensor
from . import functional as F
from .common import Dropout, Linear
from .container import LayerList
from .conv_norm_pool import LayerNorm
from .layer import Layer


def _convert_attention_mask(attn_mask, dtype):
    if attn_mask is None:
        return None
    if attn_mask.dtype == torch.bool:
        return torch.where(attn_mask, torch.zeros((), dtype=dtype, device=attn_mask.device),
                           torch.full((), -1e9 if dtype != torch.float16 else -1e4, dtype=dtype, device=attn_mask.device))
    if not attn_mask.is_floating_point():
        return (attn_mask.to(dtype) - 1.0) * 1e9
    return attn_mask.to(dtype)


class MultiHeadAttention(Layer):
    Cache = collections.namedtuple("Cache", ["k", "v"])
    StaticCache = collections.namedtuple("StaticCache", ["k", "v"])

    def __init__(self, embed_dim, num_heads, dropout=0.0, kdim=None, vdim=None, need_weights=False, weight_attr=None, bias_attr=None):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout, self.need_weights = embed_dim, num_heads, dropout, need_weights
        self.kdim, self.vdim = kdim or embed_dim, vdim or embed_dim
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.q_proj = Linear(embed_dim, embed_dim, weight_attr, bias_attr)
        self.k_proj = Linear(self.kdim, embed_dim, weight_attr, bias_attr)
        self.v_proj = Linear(self.vdim, embed_dim, weight_attr, bias_attr)
        self.out_proj = Linear(embed_dim, embed_dim, weight_attr, bias_attr)

    def _split(self, x):
        b, s, _ = x.shape
        return x.reshape([b, s, self.num_heads, self.head_dim])

    def compute_kv(self, key, value):
        return self._split(self.k_proj(key)), self._split(self.v_proj(value))

    def gen_cache(self, key, value=None, type=Cache):  # noqa: A002
        if type == MultiHeadAttention.StaticCache:
            k, v = self.compute_kv(key, value)
            return self.StaticCache(k, v)
        if value is None:
            b = key.shape[0]
            z = torch.zeros(b, 0, self.num_heads, self.head_dim, dtype=key.dtype, device=key.device).as_subclass(Tensor)
            return self.Cache(z, z)
        return self.Cache(key, value)

    def forward(self, query, key=None, value=None, attn_mask=None, cache=None):
        key = query if key is None else key
        value = query if value is None else value
        q = self._split(self.q_proj(query))
        if isinstance(cache, self.StaticCache):
            k, v = cache.k, cache.v
        else:
            k, v = self.compute_kv(key, value)
        if isinstance(cache, self.Cache):
            k = torch.cat([cache.k, k], 1)
            v = torch.cat([cache.v, v], 1)
            cache = self.Cache(k, v)
        mask = _convert_attention_mask(attn_mask, q.dtype)
        if self.need_weights:
            scores = torch.einsum("bqhd,bkhd->bhqk", q, k) * (self.head_dim ** -0.5)
            if mask is not None:
                scores = scores + mask
            w = torch.softmax(scores, -1)
            wd = F.dropout(w, self.dropout, training=self.training)
            out = torch.einsum("bhqk,bkhd->bqhd", wd, v)
        else:
            w = None
            out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=self.dropout, training=self.training)
        b, s = out.shape[0], out.shape[1]
        out = self.out_proj(out.reshape([b, s, self.embed_dim]))
        outs = [out]
        if self.need_weights:
            outs.append(w)
        if cache is not None:
            outs.append(cache)
        return out if len(outs) == 1 else tuple(outs)


def _act(name):
    return getattr(F, name)


class TransformerEncoderLayer(Layer):
    def __init__(self, d_model, nhead, dim_feedforward, dropout=0.1, activation="relu", attn_dropout=None, act_dropout=None,
                 normalize_before=False, weight_attr=None, bias_attr=None, layer_norm_eps=1e-5):
        super().__init__()
        attn_dropout = dropout if attn_dropout is None else attn_dropout
        act_dropout = dropout if act_dropout is None else act_dropout
        self.normalize_before = normalize_before
        self.self_attn = MultiHeadAttention(d_model, nhead, attn_dropout, weight_attr=weight_attr, bias_attr=bias_attr)
        self.linear1 = Linear(d_model, dim_feedforward, weight_attr, bias_attr)
        self.dropout = Dropout(act_dropout, mode="upscale_in_train")
        self.linear2 = Linear(dim_feedforward, d_model, weight_attr, bias_attr)
        self.norm1 = LayerNorm(d_model, layer_norm_eps)
        self.norm2 = LayerNorm(d_model, layer_norm_eps)
        self.dropout1 = Dropout(dropout, mode="upscale_in_train")
        self.dropout2 = Dropout(dropout, mode="upscale_in_train")
        self.activation = _act(activation)

    def forward(self, src, src_mask=None, cache=None):
        residual = src
        if self.normalize_before:
            src = self.norm1(src)
        if cache is None:
            src = self.self_attn(src, src, src, src_mask)
        else:
            src, cache = self.self_attn(src, src, src, src_mask, cache)
        src = residual + self.dropout1(src)
        if not self.normalize_before:
            src = self.norm1(src)
        residual = src
        if self.normalize_before:
            src = self.norm2(src)
        src = self.linear2(self.dropout(self.activation(self.linear1(src))))
        src = residual + self.dropout2(src)
        if not self.normalize_before:
            src = self.norm2(src)
        return src if cache is None else (src, cache)

    def gen_cache(self, src):
        return self.self_attn.gen_cache(src, type=MultiHeadAttention.Cache)


class TransformerEncoder(Layer):
    def __init__(self, encoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = LayerList([encoder_layer if i == 0 else copy.deepcopy(encoder_layer) for i in range(num_layers)])
        self.num_layers, self.norm = num_layers, norm

    def forward(self, src, src_mask=None, cache=None):
        out = src
        new_caches = []
        for i, mod in enumerate(self.layers):
            if cache is None:
                out = mod(out, src_mask)
            else:
                out, c = mod(out, src_mask, cache[i])
                new_caches.append(c)
        if self.norm is not None:
            out = self.norm(out)
        return out if cache is None else (out, new_caches)

    def gen_cache(self, src):
        return [l.gen_cache(src) for l in self.layers]


class TransformerDecoderLayer(Layer):
    def __init__(self, d_model, nhead, dim_feedforward, dropout=0.1, activation="relu", attn_dropout=None, act_dropout=None,
                 normalize_before=False, weight_attr=None, bias_attr=None, layer_norm_eps=1e-5):
        super().__init__()
        attn_dropout = dropout if attn_dropout is None else attn_dropout
        act_dropout = dropout if act_dropout is None else act_dropout
        self.normalize_before = normalize_before
        self.self_attn = MultiHeadAttention(d_model, nhead, attn_dropout, weight_attr=weight_attr, bias_attr=bias_attr)
        self.cross_attn = MultiHeadAttention(d_model, nhead, attn_dropout, weight_attr=weight_attr, bias_attr=bias_attr)
        self.linear1 = Linear(d_model, dim_feedforward, weight_attr, bias_attr)
        self.dropout = Dropout(act_dropout, mode="upscale_in_train")
        self.linear2 = Linear(dim_feedforward, d_model, weight_attr, bias_attr)
        self.norm1, self.norm2, self.norm3 = LayerNorm(d_model, layer_norm_eps), LayerNorm(d_model, layer_norm_eps), LayerNorm(d_model, layer_norm_eps)
        self.dropout1, self.dropout2, self.dropout3 = (Dropout(dropout, mode="upscale_in_train") for _ in range(3))
        self.activation = _act(activation)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, cache=None):
        residual = tgt
        if self.normalize_before:
            tgt = self.norm1(tgt)
        if cache is None:
            tgt = self.self_attn(tgt, tgt, tgt, tgt_mask, None)
        else:
            tgt, inc = self.self_attn(tgt, tgt, tgt, tgt_mask, cache[0])
        tgt = residual + self.dropout1(tgt)
        if not self.normalize_before:
            tgt = self.norm1(tgt)
        residual = tgt
        if self.normalize_before:
            tgt = self.norm2(tgt)
        if cache is None:
            tgt = self.cross_attn(tgt, memory, memory, memory_mask, None)
        else:
            tgt, static = self.cross_attn(tgt, memory, memory, memory_mask, cache[1])
        tgt = residual + self.dropout2(tgt)
        if not self.normalize_before:
            tgt = self.norm2(tgt)
        residual = tgt
        if self.normalize_before:
            tgt = self.norm3(tgt)
        tgt = self.linear2(self.dropout(self.activation(self.linear1(tgt))))
        tgt = residual + self.dropout3(tgt)
        if not self.normalize_before:
            tgt = self.norm3(tgt)
        return tgt if cache is None else (tgt, (inc, static))

    def gen_cache(self, memory):
        return (self.self_attn.gen_cache(memory, type=MultiHeadAttention.Cache),
                self.cross_attn.gen_cache(memory, memory, type=MultiHeadAttention.StaticCache))


class TransformerDecoder(Layer):
    def __init__(self, decoder_layer, num_layers, norm=None):
        super().__init__()
        self.layers = LayerList([decoder_layer if i == 0 else copy.deepcopy(decoder_layer) for i in range(num_layers)])
        self.num_layers, self.norm = num_layers, norm

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, cache=None):
        out = tgt
        new_caches = []
        for i, mod in enumerate(self.layers):
            if cache is None:
                out = mod(out, memory, tgt_mask, memory_mask)
            else:
                out, c = mod(out, memory, tgt_mask, memory_mask, cache[i])
                new_caches.append(c)
        if self.norm is not None:
            out = self.norm(out)
        return out if cache is None else (out, new_caches)

    def gen_cache(self, memory, do_zip=False):
        c = [l.gen_cache(memory) for l in self.layers]
        return list(zip(*c)) if do_zip else c


class Transformer(Layer):
    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048, dropout=0.1,
                 activation="relu", attn_dropout=None, act_dropout=None, normalize_before=False, weight_attr=None, bias_attr=None,
                 custom_encoder=None, custom_decoder=None):
        super().__init__()
        if custom_encoder is not None:
            self.encoder = custom_encoder
        else:
            el = TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, attn_dropout, act_dropout, normalize_before, weight_attr, bias_attr)
            self.encoder = TransformerEncoder(el, num_encoder_layers, LayerNorm(d_model) if normalize_before else None)
        if custom_decoder is not None:
            self.decoder = custom_decoder
        else:
            dl = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, attn_dropout, act_dropout, normalize_before, weight_attr, bias_attr)
            self.decoder = TransformerDecoder(dl, num_decoder_layers, LayerNorm(d_model) if normalize_before else None)
        self.d_model, self.nhead = d_model, nhead

    def forward(self, src, tgt, src_mask=None, tgt_mask=None, memory_mask=None):
        memory = self.encoder(src, src_mask)
        return self.decoder(tgt, memory, tgt_mask, memory_mask)

    def generate_square_subsequent_mask(self, length):
        m = torch.full((length, length), float("-inf")).triu(1)
        return m.as_subclass(Tensor)
