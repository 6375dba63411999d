"""paddle.static.nn. Parity: python/paddle/static/nn/__init__.py (31 names: fc, conv2d, batch_norm, embedding, ...).
Each builder creates its parameters once (global scope) and applies the dygraph functional, so it records into the
current Program like any other op."""
from __future__ import annotations

import torch

from ... import nn as _nn
from ...nn import functional as F
from ...nn import initializer as I
from ...nn.layer import _make_parameter
from ...tensor import Tensor


def _act(x, act):
    return getattr(F, act)(x) if act else x


def fc(x, size, num_flatten_dims=1, weight_attr=None, bias_attr=None, activation=None, name=None):
    xs = x if isinstance(x, (list, tuple)) else [x]
    out = None
    for xi in xs:
        in_dim = 1
        for s in xi.shape[num_flatten_dims:]:
            in_dim *= s
        w = _make_parameter([in_dim, size], xi.dtype, attr=weight_attr, prefix="fc_w")
        y = (xi.flatten(num_flatten_dims) if xi.dim() > num_flatten_dims + 1 else xi) @ w   # no batch-size constant on the tape
        out = y if out is None else out + y
    if bias_attr is not False:
        b = _make_parameter([size], out.dtype, attr=bias_attr, is_bias=True, prefix="fc_b")
        out = out + b
    return _act(out, activation)


def embedding(input, size, is_sparse=False, is_distributed=False, padding_idx=None, param_attr=None, dtype="float32"):
    w = _make_parameter(list(size), dtype, attr=param_attr, default_initializer=I.XavierNormal(), prefix="emb_w")
    return F.embedding(input, w, padding_idx)


def sparse_embedding(input, size, padding_idx=None, is_test=False, entry=None, table_class="MemorySparseTable", param_attr=None, dtype="float32", slot=None):
    """Embedding whose table lives on the parameter servers when the job runs in PS mode (rows are pulled in forward, gradients pushed
    in backward: fleet/ps_mode.py:DistributedEmbedding); a local embedding otherwise. Parity: static/nn/common.py:sparse_embedding."""
    from ...distributed.fleet import ps_mode

    role = ps_mode.role()
    if role is not None and not role.is_server:
        name = getattr(param_attr, "name", None) or f"sparse_embedding_{size[0]}x{size[1]}"
        key = ("ps_emb", name)
        layer = _ps_embeddings.get(key)
        if layer is None:
            layer = _ps_embeddings[key] = ps_mode.DistributedEmbedding(name, int(size[1]))
        return layer(input)
    return embedding(input, size, is_sparse=True, padding_idx=padding_idx, param_attr=param_attr, dtype=dtype)


_ps_embeddings = {}


def _conv(n, transpose, input, num_filters, filter_size, stride=1, padding=0, dilation=1, groups=1, param_attr=None, bias_attr=None, act=None,
          data_format=None, output_size=None):
    k = (filter_size,) * n if isinstance(filter_size, int) else tuple(filter_size)
    cl = data_format in ("NHWC", "NDHWC", "NLC")
    cin = input.shape[-1] if cl else input.shape[1]
    shape = [cin, num_filters // groups, *k] if transpose else [num_filters, cin // groups, *k]
    w = _make_parameter(shape, input.dtype, attr=param_attr, prefix="conv_w")
    b = None if bias_attr is False else _make_parameter([num_filters], input.dtype, attr=bias_attr, is_bias=True, prefix="conv_b")
    fmt = data_format or {1: "NCL", 2: "NCHW", 3: "NCDHW"}[n]
    if transpose:
        fn = {2: F.conv2d_transpose, 3: F.conv3d_transpose}[n]
        out = fn(input, w, b, stride, padding, 0, groups=groups, dilation=dilation, output_size=output_size, data_format=fmt)
    else:
        fn = {2: F.conv2d, 3: F.conv3d}[n]
        out = fn(input, w, b, stride, padding, dilation, groups, fmt)
    return _act(out, act)


def conv2d(input, num_filters, filter_size, stride=1, padding=0, dilation=1, groups=1, param_attr=None, bias_attr=None, use_cudnn=True, act=None, name=None, data_format="NCHW"):
    return _conv(2, False, input, num_filters, filter_size, stride, padding, dilation, groups or 1, param_attr, bias_attr, act, data_format)


def conv3d(input, num_filters, filter_size, stride=1, padding=0, dilation=1, groups=1, param_attr=None, bias_attr=None, use_cudnn=True, act=None, name=None, data_format="NCDHW"):
    return _conv(3, False, input, num_filters, filter_size, stride, padding, dilation, groups or 1, param_attr, bias_attr, act, data_format)


def conv2d_transpose(input, num_filters, output_size=None, filter_size=None, padding=0, stride=1, dilation=1, groups=1, param_attr=None, bias_attr=None, use_cudnn=True, act=None, name=None, data_format="NCHW"):
    return _conv(2, True, input, num_filters, filter_size, stride, padding, dilation, groups or 1, param_attr, bias_attr, act, data_format, output_size)


def conv3d_transpose(input, num_filters, output_size=None, filter_size=None, padding=0, stride=1, dilation=1, groups=1, param_attr=None, bias_attr=None, use_cudnn=True, act=None, name=None, data_format="NCDHW"):
    return _conv(3, True, input, num_filters, filter_size, stride, padding, dilation, groups or 1, param_attr, bias_attr, act, data_format, output_size)


def batch_norm(input, act=None, is_test=False, momentum=0.9, epsilon=1e-05, param_attr=None, bias_attr=None, data_layout="NCHW", in_place=False, name=None,
               moving_mean_name=None, moving_variance_name=None, do_model_average_for_mean_and_var=True, use_global_stats=False):
    c = input.shape[-1] if data_layout in ("NHWC", "NDHWC", "NLC") else input.shape[1]
    w = _make_parameter([c], "float32", attr=param_attr, default_initializer=I.Constant(1.0), prefix="bn_w")
    b = _make_parameter([c], "float32", attr=bias_attr, is_bias=True, prefix="bn_b")
    mean, var = torch.zeros(c).as_subclass(Tensor), torch.ones(c).as_subclass(Tensor)
    return _act(F.batch_norm(input, mean, var, w, b, not is_test, momentum, epsilon, data_layout, use_global_stats), act)


def layer_norm(input, scale=True, shift=True, begin_norm_axis=1, epsilon=1e-05, param_attr=None, bias_attr=None, act=None, name=None):
    shape = list(input.shape[begin_norm_axis:])
    w = _make_parameter(shape, input.dtype, attr=param_attr, default_initializer=I.Constant(1.0), prefix="ln_w") if scale else None
    b = _make_parameter(shape, input.dtype, attr=bias_attr, is_bias=True, prefix="ln_b") if shift else None
    return _act(F.layer_norm(input, shape, w, b, epsilon), act)


def group_norm(input, groups, epsilon=1e-05, param_attr=None, bias_attr=None, act=None, data_layout="NCHW", name=None):
    c = input.shape[1] if data_layout == "NCHW" else input.shape[-1]
    w = _make_parameter([c], input.dtype, attr=param_attr, default_initializer=I.Constant(1.0), prefix="gn_w")
    b = _make_parameter([c], input.dtype, attr=bias_attr, is_bias=True, prefix="gn_b")
    return _act(F.group_norm(input, groups, epsilon, w, b, data_layout), act)


def instance_norm(input, epsilon=1e-05, param_attr=None, bias_attr=None, name=None):
    c = input.shape[1]
    w = _make_parameter([c], input.dtype, attr=param_attr, default_initializer=I.Constant(1.0), prefix="in_w")
    b = _make_parameter([c], input.dtype, attr=bias_attr, is_bias=True, prefix="in_b")
    return F.instance_norm(input, weight=w, bias=b, eps=epsilon)


def data_norm(input, act=None, epsilon=1e-05, param_attr=None, data_layout="NCHW", in_place=False, name=None, moving_mean_name=None,
              moving_variance_name=None, do_model_average_for_mean_and_var=True, slot_dim=-1, sync_stats=False, summary_decay_rate=0.9999999,
              enable_scale_and_shift=False):
    mean = input.mean(0, keepdim=True)
    std = (input.var(0, unbiased=False, keepdim=True) + epsilon).sqrt()
    return _act((input - mean) / std, act)


def prelu(x, mode, param_attr=None, data_format="NCHW", name=None):
    n = 1 if mode == "all" else (x.shape[1] if data_format == "NCHW" else x.shape[-1])
    w = _make_parameter([n], x.dtype, attr=param_attr, default_initializer=I.Constant(0.25), prefix="prelu_w")
    return F.prelu(x, w, data_format)


def spectral_norm(weight, dim=0, power_iters=1, eps=1e-12, name=None):
    return _nn.SpectralNorm(list(weight.shape), dim, power_iters, eps)(weight)


def bilinear_tensor_product(x, y, size, act=None, name=None, param_attr=None, bias_attr=None):
    w = _make_parameter([size, x.shape[1], y.shape[1]], x.dtype, attr=param_attr, prefix="btp_w")
    b = None if bias_attr is False else _make_parameter([1, size], x.dtype, attr=bias_attr, is_bias=True, prefix="btp_b")
    return _act(F.bilinear(x, y, w, b), act)


def deform_conv2d(x, offset, mask, num_filters, filter_size, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, im2col_step=1,
                  weight_attr=None, bias_attr=None, name=None):
    from ...vision.ops import deform_conv2d as dc

    k = (filter_size,) * 2 if isinstance(filter_size, int) else tuple(filter_size)
    w = _make_parameter([num_filters, x.shape[1] // groups, *k], x.dtype, attr=weight_attr, prefix="dcn_w")
    b = None if bias_attr is False else _make_parameter([num_filters], x.dtype, attr=bias_attr, is_bias=True, prefix="dcn_b")
    return dc(x, offset, w, b, stride, padding, dilation, deformable_groups, groups, mask)


def nce(input, label, num_total_classes, sample_weight=None, param_attr=None, bias_attr=None, num_neg_samples=None, name=None, sampler="uniform",
        custom_dist=None, seed=0, is_sparse=False):
    d = input.shape[1]
    w = _make_parameter([num_total_classes, d], input.dtype, attr=param_attr, prefix="nce_w")
    b = _make_parameter([num_total_classes], input.dtype, attr=bias_attr, is_bias=True, prefix="nce_b")
    k = num_neg_samples or 10
    neg = torch.randint(0, num_total_classes, (input.shape[0], k), device=input.device)
    lab = label.reshape([-1, 1]).long()
    idx = torch.cat([lab.as_subclass(torch.Tensor), neg], 1)
    logits = torch.einsum("bd,bkd->bk", input, w[idx]) + b[idx]
    tgt = torch.zeros_like(logits)
    tgt[:, 0] = 1.0
    return F.binary_cross_entropy_with_logits(logits, tgt, reduction="none").sum(1, keepdim=True)


def row_conv(input, future_context_size, param_attr=None, act=None):
    w = _make_parameter([future_context_size + 1, input.shape[-1]], input.dtype, attr=param_attr, prefix="rowconv_w")
    T_ = input.shape[1]
    out = 0
    for i in range(future_context_size + 1):
        shifted = torch.nn.functional.pad(input[:, i:], (0, 0, 0, i))        # stays on the recorded tensor type: every step is an op of the program
        out = out + shifted * w[i]
    return _act(out if isinstance(out, Tensor) else out.as_subclass(Tensor), act)


# control flow (python-side at record time, matching dygraph semantics)
from .control_flow import case, cond, switch_case, while_loop  # noqa: F401,E402


def static_pylayer(forward_fn, inputs, backward_fn=None, name=None):
    """`forward_fn(*inputs)` with a user-defined backward.  Without `backward_fn` the forward's ops are recorded as they are; with it the call is ONE
    recorded node around an autograd function whose backward is `backward_fn(*output_grads) -> input_grads` (reference: static_pylayer.py)."""
    if backward_fn is None:
        return forward_fn(*inputs)
    from ...framework import recording

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *xs):
            with torch.no_grad():
                out = forward_fn(*[x.as_subclass(Tensor) if isinstance(x, torch.Tensor) and not isinstance(x, Tensor) else x for x in xs])
            outs = out if isinstance(out, (list, tuple)) else (out,)
            ctx.single = not isinstance(out, (list, tuple))
            res = tuple(o.as_subclass(torch.Tensor) * 1 for o in outs)
            return res[0] if ctx.single else res

        @staticmethod
        def backward(ctx, *grads):
            g = backward_fn(*[t.as_subclass(Tensor) for t in grads])
            g = g if isinstance(g, (list, tuple)) else (g,)
            return tuple(None if t is None else t.as_subclass(torch.Tensor) for t in g)

    def _static_pylayer_op(*xs):
        out = _Fn.apply(*[x.as_subclass(torch.Tensor) if isinstance(x, torch.Tensor) else x for x in xs])
        if isinstance(out, tuple):
            return tuple(o.as_subclass(Tensor) for o in out)
        return out.as_subclass(Tensor)

    _static_pylayer_op.__name__ = "static_pylayer"
    return recording.recordable(_static_pylayer_op)(*inputs)


def py_func(func, x, out, backward_func=None, skip_vars_in_backward_input=None):
    from .. import py_func as _py_func            # one recorded node that calls `func` when the program runs

    return _py_func(func, x, out, backward_func, skip_vars_in_backward_input)


from .sequence import (sequence_concat, sequence_conv, sequence_enumerate, sequence_expand, sequence_expand_as, sequence_first_step,  # noqa: F401,E402
                       sequence_last_step, sequence_pad, sequence_pool, sequence_reshape, sequence_reverse, sequence_scatter, sequence_slice,
                       sequence_softmax, sequence_unpad)
