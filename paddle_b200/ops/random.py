"""Random ops. Parity: python/paddle/tensor/random.py."""
from __future__ import annotations

import torch

from ..framework import dtype as _dt
from ._helpers import T, dev, dt, raw, shp, to_int, wrap


from ..framework.recording import recordable as _recordable  # noqa: E402

def _fd(dtype):
    return dt(dtype, _dt.default_dtype())


@_recordable
def rand(shape, dtype=None, name=None):
    return wrap(torch.rand(shp(shape), dtype=_fd(dtype), device=dev()))


@_recordable
def randn(shape, dtype=None, name=None):
    return wrap(torch.randn(shp(shape), dtype=_fd(dtype), device=dev()))


def standard_normal(shape, dtype=None, name=None):
    return randn(shape, dtype)


@_recordable
def normal(mean=0.0, std=1.0, shape=None, name=None):
    if isinstance(mean, torch.Tensor) or isinstance(std, torch.Tensor):
        m = T(mean) if isinstance(mean, torch.Tensor) else mean
        s = T(std) if isinstance(std, torch.Tensor) else std
        ref = m if isinstance(m, torch.Tensor) else s
        return ref.new_empty(ref.size()).normal_() * s + m
    return wrap(torch.empty(shp(shape if shape is not None else [1]), dtype=_dt.default_dtype(), device=dev()).normal_(float(mean), float(std)))


def normal_(x, mean=0.0, std=1.0, name=None):
    with torch.no_grad():
        torch.Tensor.normal_(x, mean, std)
    return x


@_recordable
def uniform(shape, dtype=None, min=-1.0, max=1.0, seed=0, name=None):  # noqa: A002
    g = None
    if seed:
        g = torch.Generator(device=dev()).manual_seed(int(seed))
    return wrap(torch.empty(shp(shape), dtype=_fd(dtype), device=dev()).uniform_(float(to_int(min)), float(to_int(max)), generator=g))


def uniform_(x, min=-1.0, max=1.0, seed=0, name=None):  # noqa: A002
    with torch.no_grad():
        torch.Tensor.uniform_(x, min, max)
    return x


@_recordable
def randint(low=0, high=None, shape=(1,), dtype=None, name=None):
    if high is None:
        low, high = 0, low
    return wrap(torch.randint(int(to_int(low)), int(to_int(high)), shp(shape), dtype=dt(dtype, torch.int64), device=dev()))


@_recordable
def randint_like(x, low=0, high=None, dtype=None, name=None):
    if high is None:
        low, high = 0, low
    x = T(x)
    return wrap(torch.randint(int(low), int(high), tuple(x.size()), dtype=dt(dtype, x.dtype), device=x.device))


def randperm(n, dtype="int64", name=None):
    return wrap(torch.randperm(int(to_int(n)), dtype=dt(dtype), device=dev()))


def bernoulli(x, p=None, name=None):
    return torch.bernoulli(T(x)) if p is None else torch.bernoulli(T(x), p)


def bernoulli_(x, p=0.5, name=None):
    with torch.no_grad():
        torch.Tensor.bernoulli_(x, p)
    return x


def binomial(count, prob, name=None):
    return wrap(torch.binomial(raw(count).float(), raw(prob).float()).to(torch.int64))


def poisson(x, name=None):
    return torch.poisson(T(x))


def multinomial(x, num_samples=1, replacement=False, name=None):
    return torch.multinomial(T(x), num_samples, replacement)


def exponential_(x, lam=1.0, name=None):
    with torch.no_grad():
        torch.Tensor.exponential_(x, lam)
    return x


def standard_gamma(x, name=None):
    return wrap(torch._standard_gamma(raw(x)))


def log_normal(mean=1.0, std=2.0, shape=None, dtype=None, name=None):
    return wrap(torch.empty(shp(shape if shape is not None else [1]), dtype=_fd(dtype), device=dev()).log_normal_(mean, std))


def log_normal_(x, mean=1.0, std=2.0, name=None):
    with torch.no_grad():
        torch.Tensor.log_normal_(x, mean, std)
    return x


@_recordable
def rand_like(x, dtype=None, name=None):
    return wrap(torch.rand_like(raw(x), dtype=dt(dtype)))


@_recordable
def randn_like(x, dtype=None, name=None):
    return wrap(torch.randn_like(raw(x), dtype=dt(dtype)))


__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "T", "dev", "dt", "raw", "shp", "to_int", "wrap", "annotations")]


# static programs record these as single ops (their bodies compute on raw tensors / read values; framework/recording.py)
from ..framework.recording import make_recordable as _make_recordable  # noqa: E402

# samplers: every run of a program draws anew, also when nothing of the program feeds them (constant probabilities, shapes only) - recorded unconditionally
_make_recordable(globals(), ['rand', 'randn', 'standard_normal', 'normal', 'uniform', 'randint', 'randperm', 'log_normal', 'binomial', 'standard_gamma', 'bernoulli',
                             'poisson', 'multinomial', 'randint_like', 'rand_like', 'randn_like'], always=True)
