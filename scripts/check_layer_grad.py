"""Fused vs unfused (plain PyTorch) forward/backward of one Llama-2-13B decoder layer on one GPU, bf16.
Prints output / input-gradient / weight-gradient relative errors and the gradient gain |dh| / |dout|."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import paddle_b200 as paddle  # noqa: E402
from paddle_b200.models import llama as L  # noqa: E402

paddle.set_device("gpu:0")
paddle.set_default_dtype("bfloat16")
cfg = L.llama2_13b()
cfg.num_hidden_layers = 1
cfg.max_position_embeddings = 4096
paddle.seed(0)
layer = L.LlamaDecoderLayer(cfg)


def run(h0, g, fused):
    paddle.set_flags({"FLAGS_use_fused_kernels": fused})
    for p in layer.parameters():
        p.clear_grad()
    h = h0.clone().as_subclass(paddle.Tensor)
    h.stop_gradient = False
    out = layer(h)
    out.backward(g.as_subclass(paddle.Tensor))
    gw = {n: p.grad.as_subclass(torch.Tensor).float().clone() for n, p in layer.named_parameters()}
    return out.as_subclass(torch.Tensor).float().detach(), h.grad.as_subclass(torch.Tensor).float().clone(), gw


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


for B, S, scale in ((1, 4096, 5.0), (1, 2048, 5.0), (2, 1024, 0.02)):
    torch.manual_seed(1)
    h0 = (torch.randn(B, S, cfg.hidden_size, device="cuda") * scale).bfloat16()
    g = (torch.randn(B, S, cfg.hidden_size, device="cuda") * 1e-3).bfloat16()
    of, dhf, gwf = run(h0, g, True)
    ou, dhu, gwu = run(h0, g, False)
    print(f"B={B} S={S} |h|rms={scale}: out rel {rel(of, ou):.3e}  dh rel {rel(dhf, dhu):.3e}  gain fused {float(dhf.norm() / g.float().norm()):.3f} "
          f"unfused {float(dhu.norm() / g.float().norm()):.3f}", flush=True)
    for n in gwf:
        print(f"    dW {n:40s} rel {rel(gwf[n], gwu[n]):.3e}  |fused| {float(gwf[n].norm()):.4g} |ref| {float(gwu[n].norm()):.4g}", flush=True)
