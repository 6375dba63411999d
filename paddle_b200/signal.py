"""paddle.signal. Parity: python/paddle/signal.py (stft, istft)."""
import torch

from .ops._helpers import raw, wrap


def stft(x, n_fft, hop_length=None, win_length=None, window=None, center=True, pad_mode="reflect", normalized=False, onesided=True, name=None):
    return wrap(torch.stft(raw(x), n_fft, hop_length, win_length, None if window is None else raw(window), center, pad_mode, normalized, onesided, return_complex=True))


def istft(x, n_fft, hop_length=None, win_length=None, window=None, center=True, normalized=False, onesided=True, length=None, return_complex=False, name=None):
    return wrap(torch.istft(raw(x), n_fft, hop_length, win_length, None if window is None else raw(window), center, normalized, onesided, length, return_complex))


# static programs record these as single ops (their bodies compute on raw tensors / read values; framework/recording.py)
from .framework.recording import make_recordable as _make_recordable  # noqa: E402

_make_recordable(globals(), ['stft', 'istft'])
