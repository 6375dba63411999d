"""paddle.audio. Parity: python/paddle/audio/{functional,features,backends,datasets}."""
from . import backends, datasets, features, functional  # noqa: F401
from .backends import info, load, save  # noqa: F401
