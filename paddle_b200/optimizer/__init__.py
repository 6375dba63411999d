"""paddle.optimizer. Parity: python/paddle/optimizer/__init__.py."""
from . import lr  # noqa: F401
from .optimizer import (ASGD, LBFGS, SGD, Adadelta, Adagrad, Adam, Adamax, AdamW, Lamb, Momentum, NAdam, Optimizer, RAdam, RMSProp,  # noqa: F401
                        Rprop)

__all__ = ["Optimizer", "Adagrad", "Adam", "AdamW", "Adamax", "RMSProp", "Adadelta", "SGD", "Rprop", "Momentum", "Lamb", "LBFGS",
           "NAdam", "RAdam", "ASGD"]
