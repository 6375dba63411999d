"""paddle.base.dygraph: guard / to_variable / no_grad."""
import contextlib

from ..autograd import no_grad  # noqa: F401
from ..nn.layer import Layer  # noqa: F401
from ..tensor import to_tensor


@contextlib.contextmanager
def guard(place=None):
    yield


def to_variable(value, name=None, zero_copy=None, dtype=None):
    return to_tensor(value, dtype=dtype)
