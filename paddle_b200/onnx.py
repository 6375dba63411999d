"""paddle.onnx.export. Parity: python/paddle/onnx/export.py (delegates to paddle2onnx in the reference).
Here the Layer is a torch-backed module, so export goes through torch.onnx on a thin nn.Module adapter."""
import torch


def export(layer, path, input_spec=None, opset_version=9, **configs):
    class _Adapter(torch.nn.Module):
        def __init__(self, l):
            super().__init__()
            self._l = l
            for i, p in enumerate(l.parameters()):
                self.register_parameter(f"p{i}", torch.nn.Parameter(p.as_subclass(torch.Tensor), requires_grad=False))

        def forward(self, *a):
            out = self._l(*a)
            return out.as_subclass(torch.Tensor) if isinstance(out, torch.Tensor) else out

    if not input_spec:
        raise ValueError("input_spec is required for onnx export")
    from .tensor import Tensor

    args = tuple(torch.zeros([1 if (s is None or s < 0) else s for s in spec.shape], dtype=spec.dtype).as_subclass(Tensor) if not isinstance(spec, torch.Tensor) else spec for spec in input_spec)
    layer.eval()
    try:
        torch.onnx.export(_Adapter(layer), args, path + ".onnx", opset_version=max(opset_version, 13))
    except Exception as e:  # onnx package is not part of this image
        raise RuntimeError(f"onnx export unavailable in this environment: {e}") from e
