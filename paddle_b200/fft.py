"""paddle.fft. Parity: python/paddle/fft.py."""
import torch

from .ops._helpers import T, ax


def _n(fn):
    def op(x, n=None, axis=-1, norm="backward", name=None):
        return fn(T(x), n=n, dim=axis, norm=norm)

    return op


def _nd(fn):
    def op(x, s=None, axes=None, norm="backward", name=None):
        return fn(T(x), s=s, dim=ax(axes), norm=norm)

    return op


def _2d(fn):
    def op(x, s=None, axes=(-2, -1), norm="backward", name=None):
        return fn(T(x), s=s, dim=tuple(axes), norm=norm)

    return op


fft, ifft, rfft, irfft, hfft, ihfft = map(_n, (torch.fft.fft, torch.fft.ifft, torch.fft.rfft, torch.fft.irfft, torch.fft.hfft, torch.fft.ihfft))
fftn, ifftn, rfftn, irfftn = map(_nd, (torch.fft.fftn, torch.fft.ifftn, torch.fft.rfftn, torch.fft.irfftn))
fft2, ifft2, rfft2, irfft2 = map(_2d, (torch.fft.fft2, torch.fft.ifft2, torch.fft.rfft2, torch.fft.irfft2))


def hfft2(x, s=None, axes=(-2, -1), norm="backward", name=None):
    return torch.fft.hfft2(T(x), s=s, dim=tuple(axes), norm=norm)


def ihfft2(x, s=None, axes=(-2, -1), norm="backward", name=None):
    return torch.fft.ihfft2(T(x), s=s, dim=tuple(axes), norm=norm)


def hfftn(x, s=None, axes=None, norm="backward", name=None):
    return torch.fft.hfftn(T(x), s=s, dim=ax(axes), norm=norm)


def ihfftn(x, s=None, axes=None, norm="backward", name=None):
    return torch.fft.ihfftn(T(x), s=s, dim=ax(axes), norm=norm)


def fftfreq(n, d=1.0, dtype=None, name=None):
    from .framework.dtype import convert_dtype, default_dtype
    from .ops._helpers import dev, wrap

    return wrap(torch.fft.fftfreq(n, d, dtype=convert_dtype(dtype) or default_dtype(), device=dev()))


def rfftfreq(n, d=1.0, dtype=None, name=None):
    from .framework.dtype import convert_dtype, default_dtype
    from .ops._helpers import dev, wrap

    return wrap(torch.fft.rfftfreq(n, d, dtype=convert_dtype(dtype) or default_dtype(), device=dev()))


def fftshift(x, axes=None, name=None):
    return torch.fft.fftshift(T(x), dim=ax(axes))


def ifftshift(x, axes=None, name=None):
    return torch.fft.ifftshift(T(x), dim=ax(axes))
