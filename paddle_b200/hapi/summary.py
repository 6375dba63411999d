"""paddle.summary / paddle.flops. Parity: python/paddle/hapi/model_summary.py, dynamic_flops.py."""
from __future__ import annotations

import numpy as np
import torch

from ..tensor import Tensor


def _make_inputs(input_size, dtypes, input):
    if input is not None:
        return input if isinstance(input, (list, tuple)) else [input]
    sizes = input_size if isinstance(input_size, (list,)) and input_size and isinstance(input_size[0], (list, tuple)) else [input_size]
    outs = []
    for i, s in enumerate(sizes):
        s = [1 if (d is None or d == -1) else int(d) for d in s]
        dt = (dtypes[i] if isinstance(dtypes, (list, tuple)) else dtypes) or "float32"
        from ..framework.dtype import convert_dtype

        d = convert_dtype(dt)
        outs.append((torch.rand(s) if d.is_floating_point else torch.zeros(s, dtype=d)).to(d).as_subclass(Tensor))
    return outs


def summary(net, input_size=None, dtypes=None, input=None):
    rows, hooks = [], []

    def add_hook(layer, name):
        def hook(l, inp, out):
            o = out[0] if isinstance(out, (tuple, list)) else out
            n_params = sum(p.numel() for p in l._parameters.values() if p is not None)
            rows.append((f"{type(l).__name__}-{len(rows) + 1}", list(inp[0].shape) if inp and hasattr(inp[0], "shape") else [], list(o.shape) if hasattr(o, "shape") else [], n_params))

        hooks.append(layer.register_forward_post_hook(hook))

    for name, l in net.named_sublayers():
        if not l._sub_layers:
            add_hook(l, name)
    was_training = net.training
    net.eval()
    with torch.no_grad():
        net(*_make_inputs(input_size, dtypes, input))
    if was_training:
        net.train()
    for h in hooks:
        h.remove()
    total = sum(p.numel() for p in net.parameters())
    trainable = sum(p.numel() for p in net.parameters() if not p.stop_gradient)
    line = "-" * 90
    print(line)
    print(f"{'Layer (type)':<28}{'Input Shape':<24}{'Output Shape':<24}{'Param #':>12}")
    print("=" * 90)
    for n, i, o, p in rows:
        print(f"{n:<28}{str(i):<24}{str(o):<24}{p:>12,}")
    print("=" * 90)
    print(f"Total params: {total:,}\nTrainable params: {trainable:,}\nNon-trainable params: {total - trainable:,}")
    print(line)
    return {"total_params": total, "trainable_params": trainable}


def flops(net, input_size, custom_ops=None, print_detail=False):
    """Multiply-accumulate based FLOPs estimate via forward hooks. Parity: hapi/dynamic_flops.py."""
    from .. import nn

    total = [0]
    detail, hooks = [], []

    def count(l, inp, out):
        x = inp[0]
        o = out[0] if isinstance(out, (tuple, list)) else out
        f = 0
        if custom_ops and type(l) in custom_ops:
            f = custom_ops[type(l)](l, inp, out) or 0
        elif isinstance(l, (nn.Conv1D, nn.Conv2D, nn.Conv3D, nn.Conv1DTranspose, nn.Conv2DTranspose, nn.Conv3DTranspose)):
            k = int(np.prod(l.weight.shape[2:])) * l.weight.shape[1]
            f = o.numel() * (k + (1 if l.bias is not None else 0))
        elif isinstance(l, nn.Linear):
            f = o.numel() * l.weight.shape[0] + (o.numel() if l.bias is not None else 0)
        elif isinstance(l, (nn.BatchNorm1D, nn.BatchNorm2D, nn.BatchNorm3D, nn.LayerNorm, nn.GroupNorm)):
            f = 2 * x.numel()
        elif isinstance(l, (nn.ReLU, nn.ReLU6, nn.LeakyReLU, nn.Sigmoid, nn.Tanh, nn.GELU, nn.Hardswish)):
            f = x.numel()
        elif isinstance(l, (nn.AvgPool1D, nn.AvgPool2D, nn.AvgPool3D, nn.AdaptiveAvgPool1D, nn.AdaptiveAvgPool2D, nn.AdaptiveAvgPool3D, nn.MaxPool2D)):
            f = o.numel()
        total[0] += int(f)
        detail.append((type(l).__name__, int(f)))

    for _, l in net.named_sublayers():
        if not l._sub_layers:
            hooks.append(l.register_forward_post_hook(count))
    was = net.training
    net.eval()
    with torch.no_grad():
        net(*_make_inputs(input_size, None, None))
    if was:
        net.train()
    for h in hooks:
        h.remove()
    if print_detail:
        for n, f in detail:
            print(f"{n:<24}{f:>16,}")
    print(f"Total Flops: {total[0]}     Total Params: {sum(p.numel() for p in net.parameters())}")
    return total[0]
