"""paddle.incubate.nn. Parity: python/paddle/incubate/nn/__init__.py (FusedMultiHeadAttention, FusedFeedForward,
FusedTransformerEncoderLayer, FusedMultiTransformer, FusedLinear, FusedBiasDropoutResidualLayerNorm, FusedDropoutAdd)."""
from __future__ import annotations

import torch

from ...nn import initializer as I
from ...nn.layer import Layer
from . import functional  # noqa: F401
from . import functional as FF


class FusedLinear(Layer):
    def __init__(self, in_features, out_features, weight_attr=None, bias_attr=None, transpose_weight=False, name=None):
        super().__init__()
        shape = [out_features, in_features] if transpose_weight else [in_features, out_features]
        self.weight = self.create_parameter(shape, attr=weight_attr)
        self.bias = self.create_parameter([out_features], attr=bias_attr, is_bias=True)
        self.transpose_weight = transpose_weight

    def forward(self, x):
        return FF.fused_linear(x, self.weight, self.bias, self.transpose_weight)


class FusedDropoutAdd(Layer):
    def __init__(self, p=0.5, mode="upscale_in_train", name=None):
        super().__init__()
        self.p, self.mode = p, mode

    def forward(self, x, y):
        return FF.fused_dropout_add(x, y, self.p, self.training, self.mode)


class FusedBiasDropoutResidualLayerNorm(Layer):
    def __init__(self, embed_dim, dropout_rate=0.5, weight_attr=None, bias_attr=None, epsilon=1e-05, name=None):
        super().__init__()
        self.linear_bias = self.create_parameter([embed_dim], attr=bias_attr, is_bias=True)
        self.ln_scale = self.create_parameter([embed_dim], attr=weight_attr, default_initializer=I.Constant(1.0))
        self.ln_bias = self.create_parameter([embed_dim], is_bias=True)
        self.dropout_rate, self.epsilon = dropout_rate, epsilon

    def forward(self, x, residual):
        return FF.fused_bias_dropout_residual_layer_norm(x, residual, self.linear_bias, self.ln_scale, self.ln_bias, self.dropout_rate, self.epsilon, self.training)


class FusedMultiHeadAttention(Layer):
    def __init__(self, embed_dim, num_heads, dropout_rate=0.5, attn_dropout_rate=0.5, kdim=None, vdim=None, normalize_before=False, need_weights=False,
                 qkv_weight_attr=None, qkv_bias_attr=None, linear_weight_attr=None, linear_bias_attr=None, pre_ln_scale_attr=None, pre_ln_bias_attr=None,
                 ln_scale_attr=None, ln_bias_attr=None, epsilon=1e-5, nranks=1, ring_id=-1, transpose_qkv_wb=False, name=None):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.normalize_before, self.dropout_rate, self.attn_dropout_rate, self.epsilon = normalize_before, dropout_rate, attn_dropout_rate, epsilon
        self.transpose_qkv_wb = transpose_qkv_wb
        qshape = [embed_dim, 3 * embed_dim] if transpose_qkv_wb else [3, num_heads, self.head_dim, embed_dim]
        self.qkv_weight = self.create_parameter(qshape, attr=qkv_weight_attr)
        self.qkv_bias = self.create_parameter([3 * embed_dim] if transpose_qkv_wb else [3, num_heads, self.head_dim], attr=qkv_bias_attr, is_bias=True)
        self.linear_weight = self.create_parameter([embed_dim, embed_dim], attr=linear_weight_attr)
        self.linear_bias = self.create_parameter([embed_dim], attr=linear_bias_attr, is_bias=True)
        self.pre_ln_scale = self.create_parameter([embed_dim], attr=pre_ln_scale_attr, default_initializer=I.Constant(1.0)) if normalize_before else None
        self.pre_ln_bias = self.create_parameter([embed_dim], attr=pre_ln_bias_attr, is_bias=True) if normalize_before else None
        self.ln_scale = self.create_parameter([embed_dim], attr=ln_scale_attr, default_initializer=I.Constant(1.0)) if not normalize_before else None
        self.ln_bias = self.create_parameter([embed_dim], attr=ln_bias_attr, is_bias=True) if not normalize_before else None

    def forward(self, query, key=None, value=None, attn_mask=None, cache=None):
        return FF.fused_multi_head_attention(query, self.qkv_weight, self.linear_weight, self.normalize_before, self.pre_ln_scale, self.pre_ln_bias,
                                             self.ln_scale, self.ln_bias, self.epsilon, self.qkv_bias, self.linear_bias, cache, attn_mask,
                                             self.dropout_rate, self.attn_dropout_rate, self.epsilon, self.training, num_heads=self.num_heads,
                                             transpose_qkv_wb=self.transpose_qkv_wb)


class FusedFeedForward(Layer):
    def __init__(self, d_model, dim_feedforward, dropout_rate=0.1, epsilon=1e-05, activation="relu", act_dropout_rate=None, normalize_before=False,
                 linear1_weight_attr=None, linear1_bias_attr=None, linear2_weight_attr=None, linear2_bias_attr=None, ln1_scale_attr=None, ln1_bias_attr=None,
                 ln2_scale_attr=None, ln2_bias_attr=None, nranks=1, ring_id=-1, name=None):
        super().__init__()
        self.normalize_before, self.activation, self.epsilon = normalize_before, activation, epsilon
        self.dropout_rate = dropout_rate
        self.act_dropout_rate = dropout_rate if act_dropout_rate is None else act_dropout_rate
        self.linear1_weight = self.create_parameter([d_model, dim_feedforward], attr=linear1_weight_attr)
        self.linear1_bias = self.create_parameter([dim_feedforward], attr=linear1_bias_attr, is_bias=True)
        self.linear2_weight = self.create_parameter([dim_feedforward, d_model], attr=linear2_weight_attr)
        self.linear2_bias = self.create_parameter([d_model], attr=linear2_bias_attr, is_bias=True)
        self.ln1_scale = self.create_parameter([d_model], attr=ln1_scale_attr, default_initializer=I.Constant(1.0))
        self.ln1_bias = self.create_parameter([d_model], attr=ln1_bias_attr, is_bias=True)
        self.ln2_scale = self.create_parameter([d_model], attr=ln2_scale_attr, default_initializer=I.Constant(1.0))
        self.ln2_bias = self.create_parameter([d_model], attr=ln2_bias_attr, is_bias=True)

    def forward(self, src, cache=None):
        return FF.fused_feedforward(src, self.linear1_weight, self.linear2_weight, self.linear1_bias, self.linear2_bias, self.ln1_scale, self.ln1_bias,
                                    self.ln2_scale, self.ln2_bias, self.act_dropout_rate, self.dropout_rate, self.activation, self.epsilon, self.epsilon,
                                    self.normalize_before, self.training)


class FusedTransformerEncoderLayer(Layer):
    def __init__(self, d_model, nhead, dim_feedforward, dropout_rate=0.1, activation="relu", attn_dropout_rate=None, act_dropout_rate=None,
                 normalize_before=False, weight_attr=None, bias_attr=None):
        super().__init__()
        self.fused_attn = FusedMultiHeadAttention(d_model, nhead, dropout_rate, dropout_rate if attn_dropout_rate is None else attn_dropout_rate,
                                                  normalize_before=normalize_before)
        self.ffn = FusedFeedForward(d_model, dim_feedforward, dropout_rate, activation=activation, act_dropout_rate=act_dropout_rate, normalize_before=normalize_before)

    def forward(self, src, src_mask=None, cache=None):
        out = self.fused_attn(src, attn_mask=src_mask, cache=cache)
        if cache is not None:
            out, c = out
            return self.ffn(out), c
        return self.ffn(out)


class FusedMultiTransformer(Layer):
    """Decoder stack with KV cache for serving. Parity: incubate/nn/layer/fused_transformer.py:FusedMultiTransformer
    (paddle/phi/kernels/fusion/gpu/fused_multi_transformer_op.cu.h)."""

    def __init__(self, embed_dim, num_heads, dim_feedforward, dropout_rate=0.0, activation="gelu", normalize_before=True, ln_scale_attrs=None,
                 ln_bias_attrs=None, qkv_weight_attrs=None, qkv_bias_attrs=None, linear_weight_attrs=None, linear_bias_attrs=None, ffn_ln_scale_attrs=None,
                 ffn_ln_bias_attrs=None, ffn1_weight_attrs=None, ffn1_bias_attrs=None, ffn2_weight_attrs=None, ffn2_bias_attrs=None, epsilon=1e-5,
                 residual_alpha=1.0, num_layers=-1, nranks=1, trans_qkvw=True, ring_id=-1, norm_type="layernorm", use_neox_rotary_style=False,
                 gqa_group_size=-1, name=None):
        super().__init__()
        if num_layers < 0:
            num_layers = len(qkv_weight_attrs) if isinstance(qkv_weight_attrs, (list, tuple)) else 1
        self.num_layers, self.normalize_before, self.activation, self.epsilon, self.trans_qkvw = num_layers, normalize_before, activation, epsilon, trans_qkvw
        hd = embed_dim // num_heads
        names = ["ln_scales", "ln_biases", "qkv_weights", "qkv_biases", "linear_weights", "linear_biases", "ffn_ln_scales", "ffn_ln_biases",
                 "ffn1_weights", "ffn1_biases", "ffn2_weights", "ffn2_biases"]
        for n in names:
            setattr(self, "_" + n, [])
        for i in range(num_layers):
            def add(listname, shape, is_bias=False, ones=False):
                p = self.create_parameter(shape, is_bias=is_bias, default_initializer=I.Constant(1.0) if ones else None)
                self.add_parameter(f"{listname}_{i}", p)
                getattr(self, "_" + listname).append(p)

            add("ln_scales", [embed_dim], ones=True)
            add("ln_biases", [embed_dim], True)
            add("qkv_weights", [3, num_heads, hd, embed_dim] if trans_qkvw else [embed_dim, 3, num_heads, hd])
            add("qkv_biases", [3, num_heads, hd], True)
            add("linear_weights", [embed_dim, embed_dim])
            add("linear_biases", [embed_dim], True)
            add("ffn_ln_scales", [embed_dim], ones=True)
            add("ffn_ln_biases", [embed_dim], True)
            add("ffn1_weights", [embed_dim, dim_feedforward])
            add("ffn1_biases", [dim_feedforward], True)
            add("ffn2_weights", [dim_feedforward, embed_dim])
            add("ffn2_biases", [embed_dim], True)

    def forward(self, src, attn_mask=None, caches=None, pre_caches=None, rotary_embs=None, rotary_emb_dims=0, seq_lens=None, time_step=None):
        return FF.fused_multi_transformer(src, self._ln_scales, self._ln_biases, self._qkv_weights, self._qkv_biases, self._linear_weights,
                                          self._linear_biases, self._ffn_ln_scales, self._ffn_ln_biases, self._ffn1_weights, self._ffn1_biases,
                                          self._ffn2_weights, self._ffn2_biases, pre_layer_norm=self.normalize_before, epsilon=self.epsilon,
                                          cache_kvs=caches, pre_caches=pre_caches, seq_lens=seq_lens, rotary_embs=rotary_embs, time_step=time_step,
                                          attn_mask=attn_mask, dropout_rate=0.0, rotary_emb_dims=rotary_emb_dims, activation=self.activation,
                                          training=self.training, trans_qkvw=self.trans_qkvw)


def memory_efficient_attention(query, key, value, attn_bias=None, p=0.0, scale=None, training=True):
    from ...kernels import attention as KAT

    return KAT.attention(query, key, value, attn_bias, p if training else 0.0, False, scale)
