"""paddle.vision. Parity: python/paddle/vision/__init__.py."""
from . import datasets, models, ops, transforms  # noqa: F401
from .models import *  # noqa: F401,F403

_backend = ["pil"]


def set_image_backend(backend):
    if backend not in ("pil", "cv2", "tensor"):
        raise ValueError(f"Expected backend are one of ['pil', 'cv2', 'tensor'], but got {backend}")
    _backend[0] = backend


def get_image_backend():
    return _backend[0]


def image_load(path, backend=None):
    from PIL import Image

    return Image.open(path)
