"""paddle.distribution. Parity: python/paddle/distribution/*.py (27 distributions, transforms, kl_divergence/register_kl).

Thin paddle-surface classes over torch.distributions (sampling / log_prob / entropy / kl are device-side torch ops)."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributions as D
import torch.distributions.transforms as DT

from ..tensor import Tensor


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x.as_subclass(torch.Tensor)
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.float32) if not isinstance(x, (int, float)) else float(x))
    if dtype is not None:
        t = t.to(dtype)
    if not t.is_floating_point() and not t.is_complex():
        t = t.float()
    return t


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


# ---- static programs ---------------------------------------------------------------------------------------------------------------------------
# The classes below compute through torch.distributions objects built from RAW parameter tensors, which the op tape of a static program cannot
# see.  While a program is being built, a method call on a distribution is therefore recorded as ONE node that rebuilds the distribution from its
# constructor arguments (program values among them are references, resolved at run time) and calls the method: log_prob / entropy / kl follow the feeds,
# and sample / rsample draw anew at every run.
def _enc(x):
    if isinstance(x, Distribution) and hasattr(x, "_ctor"):
        return ("__dist__", type(x), _enc(x._ctor[0]), _enc(x._ctor[1]))
    if isinstance(x, (list, tuple)):
        return type(x)(_enc(i) for i in x)
    if isinstance(x, dict):
        return {k: _enc(v) for k, v in x.items()}
    return x


def _dec(x):
    if isinstance(x, tuple) and len(x) == 4 and x[0] == "__dist__":
        return x[1](*_dec(x[2]), **_dec(x[3]))
    if isinstance(x, (list, tuple)):
        return type(x)(_dec(i) for i in x)
    if isinstance(x, dict):
        return {k: _dec(v) for k, v in x.items()}
    return x


def _replay(spec, name, margs, mkw):
    d = _dec(spec)
    attr = getattr(d, name)
    return attr(*_dec(margs), **_dec(mkw)) if callable(attr) else attr


def _replay_kl(spec_p, spec_q):
    return kl_divergence(_dec(spec_p), _dec(spec_q))


def _recording_program():
    from ..framework import recording

    return recording.current[0] if recording._inside[0] == 0 else None


def _in_program(self, name, a, k, fn):
    prog = _recording_program()
    if prog is not None and hasattr(self, "_ctor"):
        from ..framework import recording

        spec, ea, ek = _enc(self), _enc(a), _enc(k)
        sampler = name in ("sample", "rsample")
        if sampler or prog._touches_program((spec, ea, ek)):
            rec = recording.recordable(_replay, always=sampler)
            return rec(spec, name, ea, ek)
    return fn(self, *a, **k)


def _guard_method(name, fn):
    import functools

    @functools.wraps(fn)
    def method(self, *a, **k):
        return _in_program(self, name, a, k, fn)

    method._b200_guarded = True
    return method


_GUARDED = ("sample", "rsample", "log_prob", "prob", "probs", "entropy", "cdf", "icdf")


class Distribution:
    """Base class. Parity: distribution/distribution.py."""

    _d = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        init = cls.__dict__.get("__init__")
        if init is not None and not getattr(init, "_b200_captures", False):
            import functools

            @functools.wraps(init)
            def __init__(self, *a, **k):
                if not hasattr(self, "_ctor"):                     # the outermost constructor's arguments describe the object
                    self._ctor = (a, k)
                init(self, *a, **k)

            __init__._b200_captures = True
            cls.__init__ = __init__
        for name in _GUARDED:
            fn = cls.__dict__.get(name)
            if callable(fn) and not getattr(fn, "_b200_guarded", False):
                setattr(cls, name, _guard_method(name, fn))

    def __init__(self, batch_shape=(), event_shape=()):
        self._batch_shape, self._event_shape = tuple(batch_shape), tuple(event_shape)

    @property
    def batch_shape(self):
        return list(self._d.batch_shape) if self._d is not None else list(self._batch_shape)

    @property
    def event_shape(self):
        return list(self._d.event_shape) if self._d is not None else list(self._event_shape)

    @property
    def mean(self):
        return _in_program(self, "mean", (), {}, lambda s: _w(s._d.mean))

    @property
    def variance(self):
        return _in_program(self, "variance", (), {}, lambda s: _w(s._d.variance))

    @property
    def stddev(self):
        return _in_program(self, "stddev", (), {}, lambda s: _w(s._d.stddev))

    def sample(self, shape=()):
        with torch.no_grad():
            return _w(self._d.sample(tuple(shape)))

    def rsample(self, shape=()):
        return _w(self._d.rsample(tuple(shape)))

    def log_prob(self, value):
        return _w(self._d.log_prob(_t(value)))

    def prob(self, value):
        return _w(torch.exp(self._d.log_prob(_t(value))))

    probs = prob

    def entropy(self):
        return _w(self._d.entropy())

    def cdf(self, value):
        return _w(self._d.cdf(_t(value)))

    def icdf(self, value):
        return _w(self._d.icdf(_t(value)))

    def kl_divergence(self, other):
        return kl_divergence(self, other)


for _n in _GUARDED:
    if callable(Distribution.__dict__.get(_n)) and not getattr(Distribution.__dict__[_n], "_b200_guarded", False):
        setattr(Distribution, _n, _guard_method(_n, Distribution.__dict__[_n]))
Distribution.probs = Distribution.prob


class ExponentialFamily(Distribution):
    pass


class Normal(Distribution):
    def __init__(self, loc, scale, name=None):
        self.loc, self.scale = _w(_t(loc)), _w(_t(scale))
        self._d = D.Normal(_t(loc), _t(scale))


class LogNormal(Distribution):
    def __init__(self, loc, scale, name=None):
        self.loc, self.scale = _w(_t(loc)), _w(_t(scale))
        self._d = D.LogNormal(_t(loc), _t(scale))


class Uniform(Distribution):
    def __init__(self, low, high, name=None):
        self.low, self.high = _w(_t(low)), _w(_t(high))
        self._d = D.Uniform(_t(low), _t(high), validate_args=False)

    def log_prob(self, value):
        v = _t(value)
        lo, hi = _t(self.low), _t(self.high)
        inside = ((v > lo) & (v < hi)).to(v.dtype)
        return _w(torch.log(inside) - torch.log(hi - lo))


class Bernoulli(Distribution):
    def __init__(self, probs, name=None):
        self.probs_ = _w(_t(probs))
        self._d = D.Bernoulli(probs=_t(probs))

    def rsample(self, shape=(), temperature=1.0):
        return _w(D.RelaxedBernoulli(torch.as_tensor(temperature), probs=_t(self.probs_)).rsample(tuple(shape)))


class ContinuousBernoulli(Distribution):
    def __init__(self, probs, lims=(0.499, 0.501)):
        self.probs_ = _w(_t(probs))
        self._d = D.ContinuousBernoulli(probs=_t(probs), lims=lims)


class Beta(ExponentialFamily):
    def __init__(self, alpha, beta):
        self.alpha, self.beta = _w(_t(alpha)), _w(_t(beta))
        self._d = D.Beta(_t(alpha), _t(beta))


class Binomial(Distribution):
    def __init__(self, total_count, probs):
        self.total_count, self.probs_ = total_count, _w(_t(probs))
        self._d = D.Binomial(total_count=_t(total_count) if isinstance(total_count, torch.Tensor) else int(total_count), probs=_t(probs))

    def entropy(self):
        n = int(self._d.total_count.max().item())
        k = torch.arange(n + 1, dtype=torch.float32).reshape(-1, *[1] * len(self._d.batch_shape))
        lp = self._d.log_prob(k)
        return _w(-(lp.exp() * lp).sum(0))


class Categorical(Distribution):
    def __init__(self, logits, name=None):
        self.logits = _w(_t(logits))
        lg = _t(logits)
        # paddle's Categorical takes unnormalised *probabilities-like* logits: prob = logits / sum(logits) when positive
        self._d = D.Categorical(logits=lg) if bool((lg <= 0).any()) else D.Categorical(probs=lg / lg.sum(-1, keepdim=True))

    def sample(self, shape=()):
        with torch.no_grad():
            return _w(self._d.sample(tuple(shape)))

    def probs(self, value):
        p = self._d.probs
        return _w(p[..., _t(value).long()] if p.dim() == 1 else torch.gather(p, -1, _t(value).long()))

    def log_prob(self, value):
        return _w(torch.log(self.probs(value).as_subclass(torch.Tensor)))


class Cauchy(Distribution):
    def __init__(self, loc, scale, name=None):
        self.loc, self.scale = _w(_t(loc)), _w(_t(scale))
        self._d = D.Cauchy(_t(loc), _t(scale))


class Chi2(Distribution):
    def __init__(self, df):
        self.df = _w(_t(df))
        self._d = D.Chi2(_t(df))


class Dirichlet(ExponentialFamily):
    def __init__(self, concentration):
        self.concentration = _w(_t(concentration))
        self._d = D.Dirichlet(_t(concentration))


class Exponential(ExponentialFamily):
    def __init__(self, rate):
        self.rate = _w(_t(rate))
        self._d = D.Exponential(_t(rate))


class Gamma(ExponentialFamily):
    def __init__(self, concentration, rate):
        self.concentration, self.rate = _w(_t(concentration)), _w(_t(rate))
        self._d = D.Gamma(_t(concentration), _t(rate))


class Geometric(Distribution):
    def __init__(self, probs):
        self.probs_ = _w(_t(probs))
        self._d = D.Geometric(probs=_t(probs))

    def pmf(self, k):
        return self.prob(k)

    def log_pmf(self, k):
        return self.log_prob(k)


class Gumbel(Distribution):
    def __init__(self, loc, scale):
        self.loc, self.scale = _w(_t(loc)), _w(_t(scale))
        self._d = D.Gumbel(_t(loc), _t(scale))


class Laplace(Distribution):
    def __init__(self, loc, scale):
        self.loc, self.scale = _w(_t(loc)), _w(_t(scale))
        self._d = D.Laplace(_t(loc), _t(scale))


class Multinomial(Distribution):
    def __init__(self, total_count, probs):
        self.total_count, self.probs_ = int(total_count), _w(_t(probs))
        self._d = D.Multinomial(int(total_count), probs=_t(probs))

    def entropy(self):
        # Monte-Carlo free: sum over categories of binomial-marginal terms (exact for the reference's formula)
        n, p = self.total_count, self._d.probs
        lg = torch.lgamma
        k = torch.arange(n + 1, dtype=p.dtype).reshape(-1, *[1] * p.dim())
        binom_lp = lg(torch.tensor(n + 1.0)) - lg(k + 1) - lg(n - k + 1) + k * torch.log(p) + (n - k) * torch.log1p(-p)
        term = (binom_lp.exp() * lg(k + 1)).sum(0).sum(-1)
        return _w(-lg(torch.tensor(n + 1.0)) - n * (p * torch.log(p)).sum(-1) + term)


class MultivariateNormal(Distribution):
    def __init__(self, loc, covariance_matrix=None, precision_matrix=None, scale_tril=None):
        self.loc = _w(_t(loc))
        self._d = D.MultivariateNormal(_t(loc), None if covariance_matrix is None else _t(covariance_matrix),
                                       None if precision_matrix is None else _t(precision_matrix), None if scale_tril is None else _t(scale_tril))
        self.covariance_matrix = _w(self._d.covariance_matrix)
        self.scale_tril = _w(self._d.scale_tril)


class Poisson(Distribution):
    def __init__(self, rate):
        self.rate = _w(_t(rate))
        self._d = D.Poisson(_t(rate))

    def entropy(self):
        r = _t(self.rate)
        n = int(max(30, (r.max() + 10 * r.max().sqrt()).item()))
        k = torch.arange(n, dtype=r.dtype).reshape(-1, *[1] * r.dim())
        lp = self._d.log_prob(k)
        return _w(-(lp.exp() * lp).sum(0))


class StudentT(Distribution):
    def __init__(self, df, loc, scale, name=None):
        self.df, self.loc, self.scale = _w(_t(df)), _w(_t(loc)), _w(_t(scale))
        self._d = D.StudentT(_t(df), _t(loc), _t(scale))


class LKJCholesky(Distribution):
    def __init__(self, dim=2, concentration=1.0, sample_method="onion"):
        self.dim, self.concentration = dim, _w(_t(concentration))
        self._d = D.LKJCholesky(dim, _t(concentration))


class Independent(Distribution):
    def __init__(self, base, reinterpreted_batch_rank):
        self._base = base
        self._d = D.Independent(base._d, reinterpreted_batch_rank)


# ------------------------------------------------------------------------------------------------ transforms
class Transform:
    _t = None

    def forward(self, x):
        return _w(self._t(_t(x)))

    def inverse(self, y):
        return _w(self._t.inv(_t(y)))

    def forward_log_det_jacobian(self, x):
        x = _t(x)
        return _w(self._t.log_abs_det_jacobian(x, self._t(x)))

    def inverse_log_det_jacobian(self, y):
        y = _t(y)
        return _w(-self._t.log_abs_det_jacobian(self._t.inv(y), y))

    def forward_shape(self, shape):
        return tuple(self._t.forward_shape(tuple(shape)))

    def inverse_shape(self, shape):
        return tuple(self._t.inverse_shape(tuple(shape)))

    def __call__(self, x):
        if isinstance(x, Distribution):
            return TransformedDistribution(x, [self])
        if isinstance(x, Transform):
            return ChainTransform([self, x])
        return self.forward(x)


class AbsTransform(Transform):
    def __init__(self):
        self._t = DT.AbsTransform()

    def inverse(self, y):
        y = _t(y)
        return _w(-y), _w(y)


class AffineTransform(Transform):
    def __init__(self, loc, scale):
        self._t = DT.AffineTransform(_t(loc), _t(scale))


class ExpTransform(Transform):
    def __init__(self):
        self._t = DT.ExpTransform()


class PowerTransform(Transform):
    def __init__(self, power):
        self._t = DT.PowerTransform(_t(power))


class SigmoidTransform(Transform):
    def __init__(self):
        self._t = DT.SigmoidTransform()


class TanhTransform(Transform):
    def __init__(self):
        self._t = DT.TanhTransform()


class SoftmaxTransform(Transform):
    def __init__(self):
        self._t = DT.SoftmaxTransform()


class StickBreakingTransform(Transform):
    def __init__(self):
        self._t = DT.StickBreakingTransform()


class ReshapeTransform(Transform):
    def __init__(self, in_event_shape, out_event_shape):
        self._t = DT.ReshapeTransform(tuple(in_event_shape), tuple(out_event_shape))


class IndependentTransform(Transform):
    def __init__(self, base, reinterpreted_batch_rank):
        self._t = DT.IndependentTransform(base._t, reinterpreted_batch_rank)


class ChainTransform(Transform):
    def __init__(self, transforms):
        self.transforms = list(transforms)
        self._t = DT.ComposeTransform([t._t for t in self.transforms])


class StackTransform(Transform):
    def __init__(self, transforms, axis=0):
        self._t = DT.StackTransform([t._t for t in transforms], dim=axis)


class TransformedDistribution(Distribution):
    def __init__(self, base, transforms):
        self._base = base
        self._d = D.TransformedDistribution(base._d, [t._t for t in transforms])


# ------------------------------------------------------------------------------------------------ KL
_KL_REGISTRY = {}


def register_kl(cls_p, cls_q):
    def deco(fn):
        _KL_REGISTRY[(cls_p, cls_q)] = fn
        return fn

    return deco


def kl_divergence(p, q):
    prog = _recording_program()
    if prog is not None and hasattr(p, "_ctor") and hasattr(q, "_ctor"):
        sp, sq = _enc(p), _enc(q)
        if prog._touches_program((sp, sq)):
            from ..framework import recording

            return recording.recordable(_replay_kl)(sp, sq)
    for (cp, cq), fn in _KL_REGISTRY.items():
        if isinstance(p, cp) and isinstance(q, cq):
            return fn(p, q)
    return _w(D.kl_divergence(p._d, q._d))


__all__ = ["Distribution", "ExponentialFamily", "Normal", "LogNormal", "Uniform", "Bernoulli", "ContinuousBernoulli", "Beta", "Binomial",
           "Categorical", "Cauchy", "Chi2", "Dirichlet", "Exponential", "Gamma", "Geometric", "Gumbel", "Laplace", "Multinomial",
           "MultivariateNormal", "Poisson", "StudentT", "LKJCholesky", "Independent", "TransformedDistribution", "Transform", "AbsTransform",
           "AffineTransform", "ChainTransform", "ExpTransform", "IndependentTransform", "PowerTransform", "ReshapeTransform",
           "SigmoidTransform", "SoftmaxTransform", "StackTransform", "StickBreakingTransform", "TanhTransform", "kl_divergence", "register_kl"]
