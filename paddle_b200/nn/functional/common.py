"""Common functionals. Parity: python/paddle/nn/functional/common.py, input.py, extension.py, vision.py, distance.py."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from ...framework import dtype as _dt
from ...ops._helpers import T, raw, shp, to_int, wrap


from ...framework.recording import recordable as _recordable  # noqa: E402
from ...framework.recording import recordable as _recordable  # noqa: E402  (bodies that compute on raw tensors are recorded as ONE op in static programs)

def linear(x, weight, bias=None, name=None):
    """y = x @ W + b with W laid out [in, out]. Parity: nn/functional/common.py:linear.

    On CUDA bf16/fp16 this routes to the tcgen05 GEMM (``kernels.gemm``) with the bias fused in the epilogue.
    """
    from ...kernels import gemm as K

    if getattr(type(weight), "_is_dist", False) or getattr(type(x), "_is_dist", False):   # DistTensor: sharding propagation rules
        y = torch.matmul(x, weight)
        return y if bias is None else y + bias
    return K.linear(T(x), weight, bias)


@_recordable
def dropout(x, p=0.5, axis=None, training=True, mode="upscale_in_train", name=None):
    x = T(x)
    if isinstance(p, torch.Tensor):
        p = float(p.item())
    if p == 0 or (not training and mode == "upscale_in_train"):
        return x
    if not training:  # downscale_in_infer
        return x * (1.0 - p)
    if p == 1.0:
        return torch.zeros_like(x)
    if axis is None:
        out = F.dropout(x, p, True)
        return out if mode == "upscale_in_train" else out * (1.0 - p)
    axes = [axis] if isinstance(axis, int) else list(axis)
    mshape = [x.size(i) if i in [a % x.dim() for a in axes] else 1 for i in range(x.dim())]
    mask = (torch.rand(mshape, device=x.device) >= p).to(x.dtype)
    return x * mask / (1.0 - p) if mode == "upscale_in_train" else x * mask


@_recordable
def dropout2d(x, p=0.5, training=True, data_format="NCHW", name=None):
    return dropout(x, p, axis=[0, 1] if data_format == "NCHW" else [0, 3], training=training)


@_recordable
def dropout3d(x, p=0.5, training=True, data_format="NCDHW", name=None):
    return dropout(x, p, axis=[0, 1] if data_format == "NCDHW" else [0, 4], training=training)


@_recordable
def alpha_dropout(x, p=0.5, training=True, name=None):
    return F.alpha_dropout(T(x), p, training)


@_recordable
def feature_alpha_dropout(x, p=0.5, training=True, name=None):
    return F.feature_alpha_dropout(T(x), p, training)


def _pad_nd(x, pad, mode, value, data_format):
    """paddle pad order for the non-'all dims' form: [left, right, top, bottom, front, back] on the spatial dims."""
    nd = x.dim()
    channel_last = data_format in ("NLC", "NHWC", "NDHWC")
    pad = [int(to_int(p)) for p in pad]
    if mode == "constant" and len(pad) == 2 * nd:
        # pad for every dim, paddle order = dim0_lo, dim0_hi, dim1_lo, ...
        tp = []
        for i in reversed(range(nd)):
            tp += [pad[2 * i], pad[2 * i + 1]]
        return F.pad(x, tp, "constant", value)
    tmode = {"constant": "constant", "reflect": "reflect", "replicate": "replicate", "circular": "circular"}[mode]
    if channel_last:
        perm = [0, nd - 1] + list(range(1, nd - 1))
        inv = [0] + list(range(2, nd)) + [1]
        y = F.pad(x.permute(*perm), pad, tmode, **({"value": value} if tmode == "constant" else {}))
        return y.permute(*inv)
    return F.pad(x, pad, tmode, **({"value": value} if tmode == "constant" else {}))


def pad(x, pad, mode="constant", value=0.0, data_format="NCHW", pad_from_left_axis=True, name=None):
    x = T(x)
    if isinstance(pad, torch.Tensor):
        pad = pad.tolist()
    if isinstance(value, torch.Tensor):
        value = value.item()
    return _pad_nd(x, list(pad), mode, value, data_format)


def zeropad2d(x, padding, data_format="NCHW", name=None):
    return pad(x, padding, "constant", 0.0, data_format)


def embedding(x, weight, padding_idx=None, max_norm=None, norm_type=2.0, sparse=False, scale_grad_by_freq=False, name=None):
    w = T(weight)
    if padding_idx is not None and padding_idx < 0:
        padding_idx = w.size(0) + padding_idx
    return F.embedding(T(x).long(), w, padding_idx, max_norm, norm_type, scale_grad_by_freq, bool(sparse))      # sparse: row-sparse weight gradient (SelectedRows)


def one_hot(x, num_classes, name=None):
    return F.one_hot(T(x).long(), int(to_int(num_classes))).to(_dt.default_dtype())


def label_smooth(label, prior_dist=None, epsilon=0.1, name=None):
    label = T(label)
    k = label.size(-1)
    if prior_dist is not None:
        return (1 - epsilon) * label + epsilon * T(prior_dist)
    return (1 - epsilon) * label + epsilon / k


def cosine_similarity(x1, x2, axis=1, eps=1e-8):
    return F.cosine_similarity(T(x1), T(x2), dim=axis, eps=eps)


def pairwise_distance(x, y, p=2.0, epsilon=1e-6, keepdim=False, name=None):
    return F.pairwise_distance(T(x), T(y), p, epsilon, keepdim)


def pdist(x, p=2.0, name=None):
    return F.pdist(T(x), p)


def bilinear(x1, x2, weight, bias=None, name=None):
    return F.bilinear(T(x1), T(x2), T(weight), None if bias is None else T(bias).reshape(-1))


def normalize(x, p=2, axis=1, epsilon=1e-12, name=None):
    return F.normalize(T(x), p=p, dim=axis, eps=epsilon)


def _size_arg(v):
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        return [int(i) for i in v.tolist()]
    if isinstance(v, (list, tuple)):
        return [int(to_int(i)) for i in v]
    return int(v)


@_recordable
def interpolate(x, size=None, scale_factor=None, mode="nearest", align_corners=False, align_mode=0,
                data_format="NCHW", recompute_scale_factor=None, name=None):
    x = T(x)
    mode = mode.lower()
    channel_last = data_format in ("NWC", "NLC", "NHWC", "NDHWC")
    nd = x.dim()
    if channel_last:
        x = x.permute(0, nd - 1, *range(1, nd - 1))
    tmode = {"nearest": "nearest", "bilinear": "bilinear", "bicubic": "bicubic", "trilinear": "trilinear",
             "linear": "linear", "area": "area"}[mode]
    kw = {}
    if tmode in ("bilinear", "bicubic", "trilinear", "linear"):
        kw["align_corners"] = bool(align_corners)
    sf = scale_factor
    if isinstance(sf, torch.Tensor):
        sf = sf.tolist()
    if isinstance(sf, (list, tuple)):
        sf = [float(s) for s in sf]
    xr = raw(x)
    if tmode in ("bilinear", "linear", "trilinear") and not align_corners and align_mode == 1:
        # paddle align_mode=1: src = dst * scale (no half-pixel shift) -> emulate via align_corners-free 'asymmetric' grid
        out_size = _size_arg(size) if size is not None else [int(math.floor(s * (sf if not isinstance(sf, list) else sf[i]))) for i, s in enumerate(xr.shape[2:])]
        out = _asymmetric_linear(xr, out_size)
    else:
        out = F.interpolate(xr, size=_size_arg(size), scale_factor=sf, mode=tmode, recompute_scale_factor=recompute_scale_factor, **kw)
    out = wrap(out)
    if channel_last:
        out = out.permute(0, *range(2, nd), 1)
    return out


def _asymmetric_linear(x, out_size):
    for d, o in enumerate(out_size):
        dim = 2 + d
        n = x.shape[dim]
        if o == n:
            continue
        scale = n / o
        src = torch.arange(o, device=x.device, dtype=torch.float32) * scale
        i0 = src.floor().clamp(0, n - 1).long()
        i1 = (i0 + 1).clamp(max=n - 1)
        w = (src - i0.float()).to(x.dtype)
        shape = [1] * x.dim()
        shape[dim] = o
        w = w.reshape(shape)
        x = torch.index_select(x, dim, i0) * (1 - w) + torch.index_select(x, dim, i1) * w
    return x


def upsample(x, size=None, scale_factor=None, mode="nearest", align_corners=False, align_mode=0, data_format="NCHW", name=None):
    return interpolate(x, size, scale_factor, mode, align_corners, align_mode, data_format)


def unfold(x, kernel_sizes, strides=1, paddings=0, dilations=1, name=None):
    def _p(v):
        return v if isinstance(v, int) else tuple(v)

    pads = paddings
    x = T(x)
    if isinstance(pads, (list, tuple)) and len(pads) == 4:
        x = F.pad(x, [pads[1], pads[3], pads[0], pads[2]])
        pads = 0
    return F.unfold(x, _p(kernel_sizes), _p(dilations), _p(pads), _p(strides))


def fold(x, output_sizes, kernel_sizes, strides=1, paddings=0, dilations=1, name=None):
    def _p(v):
        return v if isinstance(v, int) else tuple(v)

    pads = paddings
    if isinstance(pads, (list, tuple)) and len(pads) == 4:
        pads = (pads[0], pads[1])
    return F.fold(T(x), _p(output_sizes), _p(kernel_sizes), _p(dilations), _p(pads), _p(strides))


def pixel_shuffle(x, upscale_factor, data_format="NCHW", name=None):
    x = T(x)
    if data_format == "NHWC":
        return F.pixel_shuffle(x.permute(0, 3, 1, 2), upscale_factor).permute(0, 2, 3, 1)
    return F.pixel_shuffle(x, upscale_factor)


def pixel_unshuffle(x, downscale_factor, data_format="NCHW", name=None):
    x = T(x)
    if data_format == "NHWC":
        return F.pixel_unshuffle(x.permute(0, 3, 1, 2), downscale_factor).permute(0, 2, 3, 1)
    return F.pixel_unshuffle(x, downscale_factor)


def channel_shuffle(x, groups, data_format="NCHW", name=None):
    x = T(x)
    if data_format == "NHWC":
        return F.channel_shuffle(x.permute(0, 3, 1, 2), groups).permute(0, 2, 3, 1)
    return F.channel_shuffle(x, groups)


def affine_grid(theta, out_shape, align_corners=True, name=None):
    return F.affine_grid(T(theta), shp(out_shape), align_corners=align_corners)


def grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True, name=None):
    return F.grid_sample(T(x), T(grid), mode=mode, padding_mode=padding_mode, align_corners=align_corners)


def sequence_mask(x, maxlen=None, dtype="int64", name=None):
    x = T(x)
    m = int(to_int(maxlen)) if maxlen is not None else int(x.max().item())
    return (torch.arange(m, device=x.device) < x.unsqueeze(-1)).to(_dt.convert_dtype(dtype))


def gather_tree(ids, parents):
    ids, parents = raw(ids), raw(parents)
    T_, B, W = ids.shape
    out = torch.empty_like(ids)
    out[-1] = ids[-1]
    beam = torch.arange(W, device=ids.device).expand(B, W).clone()
    beam = parents[-1]
    for t in range(T_ - 2, -1, -1):
        out[t] = torch.gather(ids[t], 1, beam)
        beam = torch.gather(parents[t], 1, beam)
    return wrap(out)


def temporal_shift(x, seg_num, shift_ratio=0.25, name=None, data_format="NCHW"):
    x = T(x)
    if data_format == "NHWC":
        x = x.permute(0, 3, 1, 2)
    nt, c, h, w = x.size()
    n = nt // seg_num
    x5 = x.reshape(n, seg_num, c, h, w)
    c1, c2 = int(c * shift_ratio), int(c * 2 * shift_ratio)
    out = torch.zeros_like(x5)
    out[:, 1:, :c1] = x5[:, :-1, :c1]
    out[:, :-1, c1:c2] = x5[:, 1:, c1:c2]
    out[:, :, c2:] = x5[:, :, c2:]
    out = out.reshape(nt, c, h, w)
    return out.permute(0, 2, 3, 1) if data_format == "NHWC" else out


def class_center_sample(label, num_classes, num_samples, group=None):
    label = raw(label)
    pos = torch.unique(label)
    if pos.numel() >= num_samples:
        sampled = pos
    else:
        mask = torch.ones(num_classes, dtype=torch.bool, device=label.device)
        mask[pos] = False
        neg = torch.nonzero(mask).reshape(-1)
        neg = neg[torch.randperm(neg.numel(), device=label.device)[: num_samples - pos.numel()]]
        sampled = torch.sort(torch.cat([pos, neg]))[0]
    remap = torch.full((num_classes,), -1, dtype=label.dtype, device=label.device)
    remap[sampled] = torch.arange(sampled.numel(), device=label.device, dtype=label.dtype)
    return wrap(remap[label]), wrap(sampled)


def diag_embed(input, offset=0, dim1=-2, dim2=-1):
    return torch.diag_embed(T(input), offset, dim1, dim2)


__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "F", "T", "raw", "shp", "to_int", "wrap", "math", "annotations")]
