"""paddle.base.core stand-ins used by user code (eager.Tensor, is_compiled_with_*, VarDesc types)."""
import torch

from ..framework.place import CPUPlace, CUDAPinnedPlace, CUDAPlace, Place, is_compiled_with_cuda, is_compiled_with_rocm, is_compiled_with_xpu  # noqa: F401
from ..tensor import Tensor


class eager:
    Tensor = Tensor


class VarDesc:
    class VarType:
        BOOL, INT8, UINT8, INT16, INT32, INT64, FP16, BF16, FP32, FP64 = (torch.bool, torch.int8, torch.uint8, torch.int16, torch.int32, torch.int64,
                                                                        torch.float16, torch.bfloat16, torch.float32, torch.float64)
        LOD_TENSOR, SELECTED_ROWS, VOCAB = 7, 8, 99


def get_cuda_device_count():
    return torch.cuda.device_count()


def is_bfloat16_supported(place=None):
    return True


def is_float16_supported(place=None):
    return True


def _get_all_register_op_kernels(lib="phi"):
    """{op: [kernel keys]} of the KernelFactory (kernels/registry.py).  Parity: core._get_all_register_op_kernels."""
    from ..kernels.registry import all_registered_kernels

    return all_registered_kernels()


def get_all_op_names():
    from ..ops import schema

    return sorted(schema.build_registry())


def __getattr__(name):
    if name in ("Scope", "_Scope"):          # legacy spelling: paddle.base.core.Scope()
        from ..static import Scope

        return Scope
    raise AttributeError(name)
