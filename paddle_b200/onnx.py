"""paddle.onnx.export. Parity: python/paddle/onnx/export.py (delegates to paddle2onnx in the reference).

The layer is traced into a Program (`jit._trace_program`) and written by the self-contained ONNX writer (`onnx_writer.py`, hand-emitted
protobuf, opset 17) — neither `onnx` nor `paddle2onnx` is required.  Ops without a converter raise NotImplementedError with the list of
supported ops."""
from . import onnx_writer


def export(layer, path, input_spec=None, opset_version=17, **configs):
    if not input_spec:
        raise ValueError("input_spec is required for onnx export")
    import torch

    from .jit import _trace_program
    from .static import InputSpec

    specs = []
    for i, spec in enumerate(input_spec):
        if isinstance(spec, InputSpec):
            specs.append(([1 if (d is None or d < 0) else int(d) for d in spec.shape], str(spec.dtype), spec.name or f"x{i}"))
        elif isinstance(spec, torch.Tensor):
            specs.append((list(torch.Tensor.size(spec)), str(spec.dtype), f"x{i}"))
        else:
            raise TypeError("input_spec entries must be InputSpec or Tensor")
    was_training = getattr(layer, "training", False)
    layer.eval()
    try:
        blob = _trace_program(layer, specs)
    finally:
        if was_training:
            layer.train()
    out = path if path.endswith(".onnx") else path + ".onnx"
    import os

    d = os.path.dirname(out)
    if d:
        os.makedirs(d, exist_ok=True)
    onnx_writer.export_program(blob, out, opset_version=max(int(opset_version), 17))
    return out
