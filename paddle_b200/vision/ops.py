"""paddle.vision.ops. Parity: python/paddle/vision/ops.py (nms, roi_align, roi_pool, psroi_pool, deform_conv2d, box_coder,
prior_box, yolo_box, yolo_loss, distribute_fpn_proposals, generate_proposals, read_file, decode_jpeg, matrix_nms)."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as TF

from ..nn.layer import Layer
from ..tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else torch.as_tensor(t)


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


def _iou(a, b):
    area_a = (a[:, 2] - a[:, 0]).clamp(min=0) * (a[:, 3] - a[:, 1]).clamp(min=0)
    area_b = (b[:, 2] - b[:, 0]).clamp(min=0) * (b[:, 3] - b[:, 1]).clamp(min=0)
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None] - inter).clamp(min=1e-10)


def _nms_single(boxes, scores, thr):
    """Greedy NMS.  The serial part (does box i survive the survivors before it) runs in the native runtime (csrc/runtime/vision_ops.cpp): CUDA boxes
    send ONE [n, n] suppression matrix to the host instead of synchronising per box; CPU boxes are scanned directly.  Python loop as the fall-back."""
    order = torch.argsort(scores, descending=True, stable=True)
    n = order.numel()
    if n == 0:
        return order
    from .._build import load

    m = load()
    sb = boxes[order].float()
    if m is not None and hasattr(m, "nms_scan"):
        if sb.is_cuda:
            keep = m.nms_scan((_iou(sb, sb) > thr).triu_(1).to(torch.uint8).cpu())
        else:
            keep = m.nms_sorted_cpu(sb.contiguous(), float(thr))
        return order[keep.to(order.device)]
    iou = _iou(sb, sb)
    keep = torch.ones(n, dtype=torch.bool, device=boxes.device)
    for i in range(n):
        if keep[i]:
            sup = iou[i] > thr
            sup[: i + 1] = False
            keep &= ~sup
    return order[keep]


def nms(boxes, iou_threshold=0.3, scores=None, category_idxs=None, categories=None, top_k=None):
    b = _raw(boxes).float()
    s = _raw(scores).float() if scores is not None else torch.arange(b.shape[0], 0, -1, device=b.device, dtype=torch.float32)
    if category_idxs is None:
        keep = _nms_single(b, s, iou_threshold)
    else:
        c = _raw(category_idxs)
        keeps = []
        for cat in (categories if categories is not None else torch.unique(c).tolist()):
            idx = torch.nonzero(c == cat).reshape(-1)
            if idx.numel():
                keeps.append(idx[_nms_single(b[idx], s[idx], iou_threshold)])
        keep = torch.cat(keeps) if keeps else torch.empty(0, dtype=torch.int64)
        keep = keep[torch.argsort(s[keep], descending=True)]
    if top_k is not None:
        keep = keep[:top_k]
    return _w(keep)


def _bilinear(feat, y, x):
    """feat [C,H,W]; y,x arbitrary-shape float coords -> [C, *shape] (zero outside [-1, size])."""
    C, H, W = feat.shape
    valid = (y >= -1) & (y <= H) & (x >= -1) & (x <= W)
    y = y.clamp(0, H - 1)
    x = x.clamp(0, W - 1)
    y0, x0 = y.floor().long(), x.floor().long()
    y1, x1 = (y0 + 1).clamp(max=H - 1), (x0 + 1).clamp(max=W - 1)
    ly, lx = y - y0.float(), x - x0.float()
    v = (feat[:, y0, x0] * ((1 - ly) * (1 - lx)) + feat[:, y0, x1] * ((1 - ly) * lx) + feat[:, y1, x0] * (ly * (1 - lx)) + feat[:, y1, x1] * (ly * lx))
    return v * valid.to(v.dtype)


def _bilinear_zero(feat, y, x):
    """Bilinear sampling with zero padding: every out-of-range corner contributes 0 (deformable-conv semantics)."""
    C, H, W = feat.shape
    y0, x0 = y.floor(), x.floor()
    ly, lx = y - y0, x - x0
    out = 0
    for dy, wy in ((0, 1 - ly), (1, ly)):
        for dx, wx in ((0, 1 - lx), (1, lx)):
            yi, xi = (y0 + dy).long(), (x0 + dx).long()
            ok = (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
            out = out + feat[:, yi.clamp(0, H - 1), xi.clamp(0, W - 1)] * (wy * wx * ok.to(feat.dtype))
    return out


def roi_align(x, boxes, boxes_num, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=True, name=None):
    """RoIAlign without a python loop over the boxes: the boxes are grouped by their sampling grid (identical (sh, sw) -> one batched gather), the
    four bilinear neighbours of every sample come from ONE index_select over the NHWC-flattened feature map per group (chunked to bound memory).
    One host read (the box sizes decide the adaptive sampling grid), no per-box synchronisation."""
    x, boxes = _raw(x), _raw(boxes).float()
    n_img, c, h, w = x.shape
    oh, ow = (output_size, output_size) if isinstance(output_size, int) else output_size
    k = boxes.shape[0]
    if k == 0:
        return _w(x.new_zeros((0, c, oh, ow)))
    img_of = torch.repeat_interleave(torch.arange(n_img, device=x.device), _raw(boxes_num).to(x.device).long())
    off = 0.5 if aligned else 0.0
    b = boxes * spatial_scale - off
    x1, y1 = b[:, 0], b[:, 1]
    rw, rh = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    if not aligned:
        rw, rh = rw.clamp(min=1.0), rh.clamp(min=1.0)
    bw, bh = rw / ow, rh / oh
    if sampling_ratio > 0:
        sh_all = torch.full((k,), int(sampling_ratio), dtype=torch.int64)
        sw_all = sh_all.clone()
    else:
        sh_all = torch.ceil(rh / oh).clamp(min=1).long().cpu()          # the one host read
        sw_all = torch.ceil(rw / ow).clamp(min=1).long().cpu()
    feat = x.float().permute(0, 2, 3, 1).reshape(n_img * h * w, c)       # NHWC rows
    out = x.new_zeros((k, c, oh, ow), dtype=torch.float32)
    key = sh_all * 100003 + sw_all
    for kv in torch.unique(key).tolist():
        idx = torch.nonzero(key == kv).reshape(-1).to(x.device)
        sh, sw = int(kv // 100003), int(kv % 100003)
        per_roi = oh * sh * ow * sw * c
        step = max(1, (1 << 24) // max(per_roi, 1))                       # <= 16 M gathered elements per corner and chunk
        for lo in range(0, idx.numel(), step):
            g = idx[lo: lo + step]
            gy = (torch.arange(oh, device=x.device).float()[None, :, None] * bh[g, None, None] + y1[g, None, None]
                  + (torch.arange(sh, device=x.device).float()[None, None, :] + 0.5) * bh[g, None, None] / sh).reshape(g.numel(), oh * sh)
            gx = (torch.arange(ow, device=x.device).float()[None, :, None] * bw[g, None, None] + x1[g, None, None]
                  + (torch.arange(sw, device=x.device).float()[None, None, :] + 0.5) * bw[g, None, None] / sw).reshape(g.numel(), ow * sw)
            yy = gy[:, :, None].expand(-1, -1, ow * sw)
            xx = gx[:, None, :].expand(-1, oh * sh, -1)
            valid = (yy >= -1.0) & (yy <= h) & (xx >= -1.0) & (xx <= w)     # samples outside the map contribute zero
            yc, xc = yy.clamp(min=0), xx.clamp(min=0)
            y0, x0 = yc.floor().clamp(max=h - 1), xc.floor().clamp(max=w - 1)
            y1i, x1i = (y0 + 1).clamp(max=h - 1), (x0 + 1).clamp(max=w - 1)
            yc = torch.where(y0 >= h - 1, y0, yc)
            xc = torch.where(x0 >= w - 1, x0, xc)
            ly, lx = yc - y0, xc - x0
            base = (img_of[g] * (h * w))[:, None, None]
            acc = 0
            for yi, xi, wt in ((y0, x0, (1 - ly) * (1 - lx)), (y0, x1i, (1 - ly) * lx), (y1i, x0, ly * (1 - lx)), (y1i, x1i, ly * lx)):
                rows = (base + yi.long() * w + xi.long()).reshape(-1)
                acc = acc + feat.index_select(0, rows).reshape(g.numel(), oh * sh, ow * sw, c) * (wt * valid)[..., None]
            out[g] = acc.reshape(g.numel(), oh, sh, ow, sw, c).mean((2, 4)).permute(0, 3, 1, 2)
    return _w(out.to(x.dtype))


def _roi_bins(lo, size, parts, limit, integer):
    """Start / end (exclusive) of the `parts` bins that split [lo, lo + size) along one axis, clipped to [0, limit].  -> two [K, parts] int64 tensors."""
    i = torch.arange(parts, device=lo.device, dtype=torch.float32)[None]
    if integer:                                            # roi_pool: integer box, bins floor / ceil of i * size / parts
        s = lo[:, None] + torch.floor(i * size[:, None] / parts)
        e = lo[:, None] + torch.ceil((i + 1) * size[:, None] / parts)
    else:                                                  # psroi_pool: real-valued box
        s = torch.floor(lo[:, None] + i * size[:, None] / parts)
        e = torch.ceil(lo[:, None] + (i + 1) * size[:, None] / parts)
    return s.clamp(0, limit).long(), e.clamp(0, limit).long()


def roi_pool(x, boxes, boxes_num, output_size, spatial_scale=1.0, name=None):
    """Max over every bin of every box, batched: bins become row / column masks and the maximum is taken in two masked reductions (columns, then
    rows) over chunks of boxes - no python loop over boxes or bins, no host read of the box coordinates."""
    x, boxes = _raw(x), _raw(boxes).float()
    oh, ow = (output_size, output_size) if isinstance(output_size, int) else output_size
    n_img, c, H, W = x.shape
    k = boxes.shape[0]
    if k == 0:
        return _w(x.new_zeros((0, c, oh, ow)))
    img_of = torch.repeat_interleave(torch.arange(n_img, device=x.device), _raw(boxes_num).to(x.device).long())
    r = torch.round(boxes * spatial_scale)
    x1, y1, x2, y2 = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
    rh, rw = (y2 - y1 + 1).clamp(min=1), (x2 - x1 + 1).clamp(min=1)
    hs, he = _roi_bins(y1, rh, oh, H, True)
    ws, we = _roi_bins(x1, rw, ow, W, True)
    ys, xs = torch.arange(H, device=x.device)[None, None], torch.arange(W, device=x.device)[None, None]
    out = x.new_zeros((k, c, oh, ow))
    neg = torch.finfo(x.dtype).min if x.dtype.is_floating_point else torch.iinfo(x.dtype).min
    step = max(1, (1 << 25) // max(c * H * W * max(ow, 1), 1))
    for lo in range(0, k, step):
        sl = slice(lo, min(k, lo + step))
        f = x[img_of[sl]]                                                        # [g, C, H, W]
        mw = (xs >= ws[sl, :, None]) & (xs < we[sl, :, None])                    # [g, ow, W]
        t = f[:, :, :, None, :].masked_fill(~mw[:, None, None], neg).amax(-1)    # [g, C, H, ow]
        mh = (ys >= hs[sl, :, None]) & (ys < he[sl, :, None])                    # [g, oh, H]
        v = t[:, :, None].masked_fill(~mh[:, None, :, :, None], neg).amax(3)     # [g, C, oh, ow]
        empty = ((he[sl] <= hs[sl])[:, :, None] | (we[sl] <= ws[sl])[:, None, :])[:, None]
        out[sl] = v.masked_fill(empty, 0)
    return _w(out)


def psroi_pool(x, boxes, boxes_num, output_size, spatial_scale=1.0, name=None):
    """Position-sensitive average pooling through a summed-area table: every bin is four gathers, whatever its size."""
    x, boxes = _raw(x), _raw(boxes).float()
    oh, ow = (output_size, output_size) if isinstance(output_size, int) else output_size
    n_img, ctot, H, W = x.shape
    C = ctot // (oh * ow)
    k = boxes.shape[0]
    if k == 0:
        return _w(x.new_zeros((0, C, oh, ow)))
    img_of = torch.repeat_interleave(torch.arange(n_img, device=x.device), _raw(boxes_num).to(x.device).long())
    b = boxes * spatial_scale
    rh, rw = (b[:, 3] - b[:, 1]).clamp(min=0.1), (b[:, 2] - b[:, 0]).clamp(min=0.1)
    hs, he = _roi_bins(b[:, 1], rh, oh, H, False)
    ws, we = _roi_bins(b[:, 0], rw, ow, W, False)
    sat = torch.nn.functional.pad(x.double().cumsum(-1).cumsum(-2), (1, 0, 1, 0))           # [N, Ctot, H+1, W+1]
    ch = (torch.arange(C, device=x.device)[:, None, None] * (oh * ow) + torch.arange(oh, device=x.device)[None, :, None] * ow
          + torch.arange(ow, device=x.device)[None, None, :])                                   # [C, oh, ow] -> input channel of every output bin
    n_i = img_of[:, None, None, None].expand(k, C, oh, ow)
    c_i = ch[None].expand(k, C, oh, ow)
    h0, h1 = hs[:, None, :, None].expand(k, C, oh, ow), he[:, None, :, None].expand(k, C, oh, ow)
    w0, w1 = ws[:, None, None, :].expand(k, C, oh, ow), we[:, None, None, :].expand(k, C, oh, ow)
    total = sat[n_i, c_i, h1, w1] - sat[n_i, c_i, h0, w1] - sat[n_i, c_i, h1, w0] + sat[n_i, c_i, h0, w0]
    area = ((h1 - h0) * (w1 - w0)).clamp(min=0)
    out = torch.where(area > 0, total / area.clamp(min=1), torch.zeros_like(total))
    return _w(out.to(x.dtype))


def deform_conv2d(x, offset, weight, bias=None, stride=1, padding=0, dilation=1, deformable_groups=1, groups=1, mask=None, name=None):
    """Deformable conv v1/v2 via bilinear sampling + grouped matmul. Parity: vision/ops.py:deform_conv2d."""
    x, offset, weight = _raw(x), _raw(offset), _raw(weight)
    N, C, H, W = x.shape
    Co, Cg, kh, kw = weight.shape
    s = (stride, stride) if isinstance(stride, int) else tuple(stride)
    p = (padding, padding) if isinstance(padding, int) else tuple(padding)
    d = (dilation, dilation) if isinstance(dilation, int) else tuple(dilation)
    Ho = (H + 2 * p[0] - d[0] * (kh - 1) - 1) // s[0] + 1
    Wo = (W + 2 * p[1] - d[1] * (kw - 1) - 1) // s[1] + 1
    base_y = (torch.arange(Ho, device=x.device) * s[0] - p[0]).float()[:, None]
    base_x = (torch.arange(Wo, device=x.device) * s[1] - p[1]).float()[None, :]
    cpg = C // deformable_groups
    K = kh * kw
    off = offset.float().reshape(N, deformable_groups, K, 2, Ho, Wo)
    ky = (torch.arange(K, device=x.device) // kw).float()[:, None, None] * d[0]
    kx = (torch.arange(K, device=x.device) % kw).float()[:, None, None] * d[1]
    yy = base_y[None] + ky + off[:, :, :, 0]                                  # [N, G, K, Ho, Wo]: every sampling position at once
    xx = base_x[None] + kx + off[:, :, :, 1]
    y0, x0 = yy.floor(), xx.floor()
    ly, lx = yy - y0, xx - x0
    feat = x.float().reshape(N * deformable_groups, cpg, H * W)
    cols = 0
    for dy, wy in ((0, 1 - ly), (1, ly)):                                        # four bilinear corners, each ONE gather; out-of-range corners add 0
        for dx, wx in ((0, 1 - lx), (1, lx)):
            yi, xi = (y0 + dy).long(), (x0 + dx).long()
            ok = (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).reshape(N * deformable_groups, 1, K * Ho * Wo).expand(-1, cpg, -1)
            wgt = (wy * wx * ok).reshape(N * deformable_groups, 1, K * Ho * Wo)
            cols = cols + feat.gather(2, idx) * wgt
    cols = cols.reshape(N, deformable_groups, cpg, K, Ho, Wo)
    if mask is not None:
        cols = cols * _raw(mask).float().reshape(N, deformable_groups, 1, K, Ho, Wo)
    cols = cols.reshape(N, C, K, Ho, Wo)
    cols = cols.reshape(N, groups, (C // groups) * kh * kw, Ho * Wo)
    wg = weight.float().reshape(groups, Co // groups, Cg * kh * kw)
    out = torch.einsum("gok,ngkl->ngol", wg, cols).reshape(N, Co, Ho, Wo)
    if bias is not None:
        out = out + _raw(bias).float().reshape(1, -1, 1, 1)
    return _w(out.to(x.dtype))


class DeformConv2D(Layer):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, deformable_groups=1, groups=1, weight_attr=None, bias_attr=None):
        super().__init__()
        k = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self._cfg = (stride, padding, dilation, deformable_groups, groups)
        self.weight = self.create_parameter([out_channels, in_channels // groups, *k], attr=weight_attr)
        self.bias = self.create_parameter([out_channels], attr=bias_attr, is_bias=True)

    def forward(self, x, offset, mask=None):
        s, p, d, dg, g = self._cfg
        return deform_conv2d(x, offset, self.weight, self.bias, s, p, d, dg, g, mask)


class RoIAlign(Layer):
    def __init__(self, output_size, spatial_scale=1.0):
        super().__init__()
        self.output_size, self.spatial_scale = output_size, spatial_scale

    def forward(self, x, boxes, boxes_num, aligned=True):
        return roi_align(x, boxes, boxes_num, self.output_size, self.spatial_scale, aligned=aligned)


class RoIPool(Layer):
    def __init__(self, output_size, spatial_scale=1.0):
        super().__init__()
        self.output_size, self.spatial_scale = output_size, spatial_scale

    def forward(self, x, boxes, boxes_num):
        return roi_pool(x, boxes, boxes_num, self.output_size, self.spatial_scale)


class PSRoIPool(Layer):
    def __init__(self, output_size, spatial_scale=1.0):
        super().__init__()
        self.output_size, self.spatial_scale = output_size, spatial_scale

    def forward(self, x, boxes, boxes_num):
        return psroi_pool(x, boxes, boxes_num, self.output_size, self.spatial_scale)


def box_coder(prior_box, prior_box_var, target_box, code_type="encode_center_size", box_normalized=True, axis=0, name=None):
    pb, tb = _raw(prior_box).float(), _raw(target_box).float()
    var = None if prior_box_var is None else (_raw(prior_box_var).float() if not isinstance(prior_box_var, (list, tuple)) else torch.tensor(prior_box_var, dtype=torch.float32, device=pb.device))
    off = 0.0 if box_normalized else 1.0
    pw, ph = pb[:, 2] - pb[:, 0] + off, pb[:, 3] - pb[:, 1] + off
    pcx, pcy = pb[:, 0] + pw * 0.5, pb[:, 1] + ph * 0.5
    if code_type == "encode_center_size":
        tw, th = tb[:, 2] - tb[:, 0] + off, tb[:, 3] - tb[:, 1] + off
        tcx, tcy = tb[:, 0] + tw * 0.5, tb[:, 1] + th * 0.5
        out = torch.stack([(tcx[:, None] - pcx[None]) / pw[None], (tcy[:, None] - pcy[None]) / ph[None],
                           torch.log((tw[:, None] / pw[None]).abs()), torch.log((th[:, None] / ph[None]).abs())], -1)
        if var is not None:
            out = out / (var if var.dim() == 1 else var[None])
        return _w(out)
    if axis == 0:
        pw, ph, pcx, pcy = pw[None], ph[None], pcx[None], pcy[None]
        v = 1.0 if var is None else (var if var.dim() == 1 else var[None])
    else:
        pw, ph, pcx, pcy = pw[:, None], ph[:, None], pcx[:, None], pcy[:, None]
        v = 1.0 if var is None else (var if var.dim() == 1 else var[:, None])
    t = tb * v
    cx, cy = t[..., 0] * pw + pcx, t[..., 1] * ph + pcy
    w, h = torch.exp(t[..., 2]) * pw, torch.exp(t[..., 3]) * ph
    return _w(torch.stack([cx - w * 0.5, cy - h * 0.5, cx + w * 0.5 - off, cy + h * 0.5 - off], -1))


def prior_box(input, image, min_sizes, max_sizes=None, aspect_ratios=(1.0,), variance=(0.1, 0.1, 0.2, 0.2), flip=False, clip=False, steps=(0.0, 0.0),
              offset=0.5, min_max_aspect_ratios_order=False, name=None):
    fh, fw = _raw(input).shape[-2:]
    ih, iw = _raw(image).shape[-2:]
    sw = steps[0] or iw / fw
    sh = steps[1] or ih / fh
    ars = [1.0]
    for a in aspect_ratios:
        if all(abs(a - e) > 1e-6 for e in ars):
            ars.append(a)
            if flip:
                ars.append(1.0 / a)
    # the (w, h) of the priors of ONE cell, in the reference's order; the grid of centres is added by broadcasting (no loop over cells)
    whs = []
    for k, ms in enumerate(min_sizes):
        if min_max_aspect_ratios_order:
            whs.append((ms, ms))
            if max_sizes:
                s_ = math.sqrt(ms * max_sizes[k])
                whs.append((s_, s_))
            whs += [(ms * math.sqrt(a), ms / math.sqrt(a)) for a in ars if abs(a - 1.0) > 1e-6]
        else:
            whs += [(ms * math.sqrt(a), ms / math.sqrt(a)) for a in ars]
            if max_sizes:
                s_ = math.sqrt(ms * max_sizes[k])
                whs.append((s_, s_))
    wh = torch.tensor(whs, dtype=torch.float64)                                        # [P, 2]
    cx = ((torch.arange(fw, dtype=torch.float64) + offset) * sw)[None, :, None]        # [1, fw, 1]
    cy = ((torch.arange(fh, dtype=torch.float64) + offset) * sh)[:, None, None]        # [fh, 1, 1]
    half_w, half_h = wh[:, 0] / 2, wh[:, 1] / 2
    boxes = torch.stack([((cx - half_w) / iw).expand(fh, fw, -1), ((cy - half_h) / ih).expand(fh, fw, -1),
                         ((cx + half_w) / iw).expand(fh, fw, -1), ((cy + half_h) / ih).expand(fh, fw, -1)], -1)
    b = boxes.to(torch.float32).reshape(fh, fw, -1, 4)
    if clip:
        b = b.clamp(0, 1)
    v = torch.tensor(variance, dtype=torch.float32).expand_as(b).clone()
    return _w(b), _w(v)


def yolo_box(x, img_size, anchors, class_num, conf_thresh, downsample_ratio, clip_bbox=True, name=None, scale_x_y=1.0, iou_aware=False, iou_aware_factor=0.5):
    x = _raw(x).float()
    N, _, H, W = x.shape
    na = len(anchors) // 2
    if iou_aware:
        ioup = torch.sigmoid(x[:, :na].reshape(N, na, 1, H, W))
        x = x[:, na:]
    x = x.reshape(N, na, 5 + class_num, H, W)
    gy, gx = torch.meshgrid(torch.arange(H, device=x.device).float(), torch.arange(W, device=x.device).float(), indexing="ij")
    bias = -0.5 * (scale_x_y - 1.0)
    bx = (torch.sigmoid(x[:, :, 0]) * scale_x_y + bias + gx) / W
    by = (torch.sigmoid(x[:, :, 1]) * scale_x_y + bias + gy) / H
    aw = torch.tensor(anchors[0::2], device=x.device).float().reshape(1, na, 1, 1)
    ah = torch.tensor(anchors[1::2], device=x.device).float().reshape(1, na, 1, 1)
    bw = torch.exp(x[:, :, 2]) * aw / (W * downsample_ratio)
    bh = torch.exp(x[:, :, 3]) * ah / (H * downsample_ratio)
    conf = torch.sigmoid(x[:, :, 4])
    if iou_aware:
        conf = conf ** (1 - iou_aware_factor) * ioup[:, :, 0] ** iou_aware_factor
    cls = torch.sigmoid(x[:, :, 5:]) * conf[:, :, None]
    isz = _raw(img_size).float()
    ih, iw = isz[:, 0].reshape(N, 1, 1, 1), isz[:, 1].reshape(N, 1, 1, 1)
    x1, y1, x2, y2 = (bx - bw / 2) * iw, (by - bh / 2) * ih, (bx + bw / 2) * iw, (by + bh / 2) * ih
    if clip_bbox:
        x1, y1 = x1.clamp(min=0), y1.clamp(min=0)
        x2, y2 = torch.min(x2, iw - 1), torch.min(y2, ih - 1)
    keep = (conf >= conf_thresh).float()
    boxes = torch.stack([x1, y1, x2, y2], -1) * keep[..., None]
    scores = cls.permute(0, 1, 3, 4, 2) * keep[..., None]
    return _w(boxes.reshape(N, -1, 4)), _w(scores.reshape(N, -1, class_num))


def yolo_loss(x, gt_box, gt_label, anchors, anchor_mask, class_num, ignore_thresh, downsample_ratio, gt_score=None, use_label_smooth=True, name=None, scale_x_y=1.0):
    """YOLOv3 loss (per-image sum). Parity: vision/ops.py:yolo_loss (yolov3_loss kernel)."""
    x = _raw(x).float()
    gtb, gtl = _raw(gt_box).float(), _raw(gt_label).long()
    N, _, H, W = x.shape
    na = len(anchor_mask)
    x = x.reshape(N, na, 5 + class_num, H, W)
    inp = H * downsample_ratio
    all_a = torch.tensor(anchors, dtype=torch.float32, device=x.device).reshape(-1, 2)
    loss = x.new_zeros(N)
    gy, gx = torch.meshgrid(torch.arange(H, device=x.device).float(), torch.arange(W, device=x.device).float(), indexing="ij")
    bias = -0.5 * (scale_x_y - 1.0)
    for n in range(N):
        px = (torch.sigmoid(x[n, :, 0]) * scale_x_y + bias + gx) / W
        py = (torch.sigmoid(x[n, :, 1]) * scale_x_y + bias + gy) / H
        ma = all_a[anchor_mask]
        pw = torch.exp(x[n, :, 2]) * ma[:, 0].reshape(na, 1, 1) / inp
        ph = torch.exp(x[n, :, 3]) * ma[:, 1].reshape(na, 1, 1) / inp
        pred = torch.stack([px - pw / 2, py - ph / 2, px + pw / 2, py + ph / 2], -1).reshape(-1, 4)
        valid = (gtb[n, :, 2] > 0) & (gtb[n, :, 3] > 0)
        g = gtb[n][valid]
        obj_mask = torch.zeros(na, H, W, device=x.device)
        noobj = torch.ones(na, H, W, device=x.device)
        if g.numel():
            gxy = torch.stack([g[:, 0] - g[:, 2] / 2, g[:, 1] - g[:, 3] / 2, g[:, 0] + g[:, 2] / 2, g[:, 1] + g[:, 3] / 2], -1)
            noobj = (( _iou(pred, gxy).max(1).values.reshape(na, H, W)) <= ignore_thresh).float()
            for k in range(g.shape[0]):
                wh = g[k, 2:4] * inp
                inter = torch.min(all_a[:, 0], wh[0]) * torch.min(all_a[:, 1], wh[1])
                best = int((inter / (all_a[:, 0] * all_a[:, 1] + wh[0] * wh[1] - inter)).argmax())
                if best not in anchor_mask:
                    continue
                a = anchor_mask.index(best)
                gi, gj = min(int(g[k, 0] * W), W - 1), min(int(g[k, 1] * H), H - 1)
                score = 1.0 if gt_score is None else float(_raw(gt_score)[n][valid][k])
                scale = (2.0 - g[k, 2] * g[k, 3]) * score
                tx, ty = g[k, 0] * W - gi, g[k, 1] * H - gj
                tw, th = torch.log(wh[0] / all_a[best, 0]), torch.log(wh[1] / all_a[best, 1])
                loss[n] += scale * (TF.binary_cross_entropy_with_logits(x[n, a, 0, gj, gi], tx) + TF.binary_cross_entropy_with_logits(x[n, a, 1, gj, gi], ty))
                loss[n] += scale * ((x[n, a, 2, gj, gi] - tw).abs() + (x[n, a, 3, gj, gi] - th).abs())
                obj_mask[a, gj, gi] = score
                noobj[a, gj, gi] = 0
                tgt = torch.full((class_num,), (1.0 / class_num if use_label_smooth else 0.0) * (1 if use_label_smooth else 0), device=x.device)
                pos = 1.0 - (1.0 / class_num if use_label_smooth else 0.0) * (1 if use_label_smooth else 0) + (1.0 / class_num if use_label_smooth else 0.0) * 0
                tgt[int(gtl[n][valid][k])] = pos if use_label_smooth else 1.0
                loss[n] += score * TF.binary_cross_entropy_with_logits(x[n, a, 5:, gj, gi], tgt, reduction="sum")
        conf = x[n, :, 4]
        loss[n] += (TF.binary_cross_entropy_with_logits(conf, torch.ones_like(conf), reduction="none") * obj_mask).sum()
        loss[n] += (TF.binary_cross_entropy_with_logits(conf, torch.zeros_like(conf), reduction="none") * noobj * (obj_mask == 0).float()).sum()
    return _w(loss)


def distribute_fpn_proposals(fpn_rois, min_level, max_level, refer_level, refer_scale, pixel_offset=False, rois_num=None, name=None):
    rois = _raw(fpn_rois).float()
    off = 1.0 if pixel_offset else 0.0
    area = (rois[:, 2] - rois[:, 0] + off) * (rois[:, 3] - rois[:, 1] + off)
    lvl = torch.floor(torch.log2(torch.sqrt(area.clamp(min=1e-6)) / refer_scale + 1e-8) + refer_level).clamp(min_level, max_level).long()
    outs, idxs, nums = [], [], []
    for l in range(min_level, max_level + 1):
        idx = torch.nonzero(lvl == l).reshape(-1)
        outs.append(_w(rois[idx]))
        idxs.append(idx)
        if rois_num is not None:
            bounds = torch.cumsum(_raw(rois_num), 0)
            img = torch.bucketize(idx, bounds, right=True)
            nums.append(_w(torch.bincount(img, minlength=bounds.numel()).to(torch.int32)))
    order = torch.cat(idxs)
    restore = torch.empty_like(order)
    restore[order] = torch.arange(order.numel(), device=order.device)
    return (outs, _w(restore.reshape(-1, 1)), nums) if rois_num is not None else (outs, _w(restore.reshape(-1, 1)))


def generate_proposals(scores, bbox_deltas, img_size, anchors, variances, pre_nms_top_n=6000, post_nms_top_n=1000, nms_thresh=0.5, min_size=0.1,
                       eta=1.0, pixel_offset=False, return_rois_num=False, name=None):
    sc, bd = _raw(scores).float(), _raw(bbox_deltas).float()
    an, va = _raw(anchors).float().reshape(-1, 4), _raw(variances).float().reshape(-1, 4)
    isz = _raw(img_size).float()
    N = sc.shape[0]
    off = 1.0 if pixel_offset else 0.0
    rois, probs, nums = [], [], []
    for n in range(N):
        s = sc[n].permute(1, 2, 0).reshape(-1)
        d = bd[n].permute(1, 2, 0).reshape(-1, 4)
        k = min(pre_nms_top_n, s.numel()) if pre_nms_top_n > 0 else s.numel()
        s, idx = s.topk(k)
        d, a, v = d[idx], an[idx], va[idx]
        aw, ah = a[:, 2] - a[:, 0] + off, a[:, 3] - a[:, 1] + off
        acx, acy = a[:, 0] + 0.5 * aw, a[:, 1] + 0.5 * ah
        cx, cy = v[:, 0] * d[:, 0] * aw + acx, v[:, 1] * d[:, 1] * ah + acy
        w = torch.exp((v[:, 2] * d[:, 2]).clamp(max=math.log(1000.0 / 16))) * aw
        h = torch.exp((v[:, 3] * d[:, 3]).clamp(max=math.log(1000.0 / 16))) * ah
        box = torch.stack([cx - w / 2, cy - h / 2, cx + w / 2 - off, cy + h / 2 - off], -1)
        ih, iw = isz[n, 0], isz[n, 1]
        box[:, 0::2] = box[:, 0::2].clamp(min=0).clamp(max=iw - off)
        box[:, 1::2] = box[:, 1::2].clamp(min=0).clamp(max=ih - off)
        keep = ((box[:, 2] - box[:, 0] + off) >= min_size) & ((box[:, 3] - box[:, 1] + off) >= min_size)
        box, s = box[keep], s[keep]
        if box.numel():
            kk = _nms_single(box, s, nms_thresh)[:post_nms_top_n]
            box, s = box[kk], s[kk]
        rois.append(box)
        probs.append(s.reshape(-1, 1))
        nums.append(box.shape[0])
    out = (_w(torch.cat(rois)), _w(torch.cat(probs)))
    return out + (_w(torch.tensor(nums, dtype=torch.int32)),) if return_rois_num else out


def matrix_nms(bboxes, scores, score_threshold, post_threshold, nms_top_k, keep_top_k, use_gaussian=False, gaussian_sigma=2.0, background_label=0,
               normalized=True, return_index=False, return_rois_num=True, name=None):
    bb, sc = _raw(bboxes).float(), _raw(scores).float()
    outs, idxs, nums = [], [], []
    for n in range(bb.shape[0]):
        dets = []
        for c in range(sc.shape[1]):
            if c == background_label:
                continue
            s = sc[n, c]
            m = s > score_threshold
            if not bool(m.any()):
                continue
            ii = torch.nonzero(m).reshape(-1)
            s, order = s[ii].sort(descending=True)
            ii = ii[order][: nms_top_k if nms_top_k > 0 else None]
            s = s[: ii.numel()]
            b = bb[n, ii]
            iou = _iou(b, b).triu(1)
            cmax = iou.max(0).values
            decay = (torch.exp(-(iou ** 2 - cmax[:, None] ** 2) / gaussian_sigma) if use_gaussian else (1 - iou) / (1 - cmax[:, None])).min(0).values
            s2 = s * decay
            k = s2 > post_threshold
            for j in torch.nonzero(k).reshape(-1).tolist():
                dets.append((float(s2[j]), c, b[j], int(ii[j]) + n * bb.shape[1]))
        dets.sort(key=lambda t: -t[0])
        dets = dets[: keep_top_k if keep_top_k > 0 else None]
        nums.append(len(dets))
        for s_, c, b, i in dets:
            outs.append(torch.cat([torch.tensor([float(c), s_]), b.cpu()]))
            idxs.append(i)
    out = torch.stack(outs) if outs else torch.zeros((0, 6))
    res = [_w(out)]
    if return_index:
        res.append(_w(torch.tensor(idxs, dtype=torch.int64).reshape(-1, 1)))
    if return_rois_num:
        res.append(_w(torch.tensor(nums, dtype=torch.int32)))
    return tuple(res) if len(res) > 1 else res[0]


def read_file(filename, name=None):
    with open(filename, "rb") as f:
        return _w(torch.frombuffer(bytearray(f.read()), dtype=torch.uint8))


def decode_jpeg(x, mode="unchanged", name=None):
    import io

    from PIL import Image

    img = Image.open(io.BytesIO(bytes(_raw(x).cpu().numpy().tobytes())))
    if mode == "gray":
        img = img.convert("L")
    elif mode == "rgb":
        img = img.convert("RGB")
    a = np.asarray(img)
    a = a[None] if a.ndim == 2 else a.transpose(2, 0, 1)
    return _w(torch.from_numpy(np.ascontiguousarray(a)))


# static programs record these as single ops (their bodies compute on raw tensors; framework/recording.py)
from ..framework.recording import make_recordable as _make_recordable  # noqa: E402

_make_recordable(globals(), ['roi_align', 'roi_pool', 'psroi_pool', 'deform_conv2d', 'box_coder', 'yolo_box', 'prior_box'])
