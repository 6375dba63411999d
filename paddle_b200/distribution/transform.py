"""paddle.distribution.transform. Parity: python/paddle/distribution/transform.py."""
from . import (AbsTransform, AffineTransform, ChainTransform, ExpTransform, IndependentTransform, PowerTransform, ReshapeTransform,  # noqa: F401
               SigmoidTransform, SoftmaxTransform, StackTransform, StickBreakingTransform, TanhTransform, Transform)
