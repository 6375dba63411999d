"""Parity: python/paddle/audio/functional/{functional,window}.py."""
from __future__ import annotations

import math

import numpy as np
import torch

from ..tensor import Tensor


def _w(t):
    return t.as_subclass(Tensor) if not isinstance(t, Tensor) else t


def hz_to_mel(freq, htk=False):
    scalar = not isinstance(freq, torch.Tensor)
    f = torch.as_tensor(freq, dtype=torch.float64) if scalar else freq.as_subclass(torch.Tensor).double()
    if htk:
        m = 2595.0 * torch.log10(1.0 + f / 700.0)
    else:
        f_sp = 200.0 / 3
        m = f / f_sp
        min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
        min_log_mel = min_log_hz / f_sp
        m = torch.where(f >= min_log_hz, min_log_mel + torch.log(f.clamp(min=1e-10) / min_log_hz) / logstep, m)
    return float(m) if scalar else _w(m.float())


def mel_to_hz(mel, htk=False):
    scalar = not isinstance(mel, torch.Tensor)
    m = torch.as_tensor(mel, dtype=torch.float64) if scalar else mel.as_subclass(torch.Tensor).double()
    if htk:
        f = 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    else:
        f_sp = 200.0 / 3
        f = f_sp * m
        min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
        min_log_mel = min_log_hz / f_sp
        f = torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), f)
    return float(f) if scalar else _w(f.float())


def mel_frequencies(n_mels=64, f_min=0.0, f_max=11025.0, htk=False, dtype="float32"):
    mels = torch.linspace(hz_to_mel(f_min, htk), hz_to_mel(f_max, htk), n_mels, dtype=torch.float64)
    return mel_to_hz(mels.as_subclass(Tensor), htk)


def fft_frequencies(sr, n_fft, dtype="float32"):
    return _w(torch.linspace(0, float(sr) / 2, 1 + n_fft // 2))


def compute_fbank_matrix(sr, n_fft, n_mels=64, f_min=0.0, f_max=None, htk=False, norm="slaney", dtype="float32"):
    f_max = f_max or float(sr) / 2
    fftfreqs = fft_frequencies(sr, n_fft).as_subclass(torch.Tensor).double()
    mel_f = mel_frequencies(n_mels + 2, f_min, f_max, htk).as_subclass(torch.Tensor).double()
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fftfreqs[None]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = torch.clamp(torch.min(lower, upper), min=0)
    if norm == "slaney":
        w = w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return _w(w.float())


def power_to_db(spect, ref_value=1.0, amin=1e-10, top_db=80.0):
    s = spect.as_subclass(torch.Tensor)
    db = 10.0 * torch.log10(torch.clamp(s, min=amin)) - 10.0 * math.log10(max(amin, ref_value))
    if top_db is not None:
        db = torch.clamp(db, min=db.max().item() - top_db)
    return _w(db)


def create_dct(n_mfcc, n_mels, norm="ortho", dtype="float32"):
    n = torch.arange(n_mels, dtype=torch.float64)
    k = torch.arange(n_mfcc, dtype=torch.float64)[:, None]
    dct = torch.cos(math.pi / n_mels * (n + 0.5) * k)
    if norm is None:
        dct = dct * 2.0
    else:
        dct[0] *= 1.0 / math.sqrt(2.0)
        dct = dct * math.sqrt(2.0 / n_mels)
    return _w(dct.t().float())


def get_window(window, win_length, fftbins=True, dtype="float64"):
    name, args = (window, ()) if isinstance(window, str) else (window[0], window[1:])
    M = win_length + (1 if fftbins else 0)
    n = torch.arange(M, dtype=torch.float64)
    if name in ("hann", "hanning"):
        w = 0.5 - 0.5 * torch.cos(2 * math.pi * n / (M - 1))
    elif name == "hamming":
        w = 0.54 - 0.46 * torch.cos(2 * math.pi * n / (M - 1))
    elif name == "blackman":
        w = 0.42 - 0.5 * torch.cos(2 * math.pi * n / (M - 1)) + 0.08 * torch.cos(4 * math.pi * n / (M - 1))
    elif name in ("rect", "boxcar", "ones"):
        w = torch.ones(M, dtype=torch.float64)
    elif name == "triang":
        w = 1 - torch.abs((n - (M - 1) / 2) / ((M + (M % 2)) / 2) if M % 2 else (n - (M - 1) / 2) / (M / 2))
    elif name == "bartlett":
        w = 1 - torch.abs(2 * n / (M - 1) - 1)
    elif name == "cosine":
        w = torch.sin(math.pi / M * (n + 0.5))
    elif name == "gaussian":
        std = args[0]
        w = torch.exp(-0.5 * ((n - (M - 1) / 2) / std) ** 2)
    elif name == "exponential":
        center, tau = (args + (None, 1.0))[:2] if args else (None, 1.0)
        center = (M - 1) / 2 if center is None else center
        w = torch.exp(-torch.abs(n - center) / tau)
    elif name == "bohman":
        x = torch.abs(torch.linspace(-1, 1, M, dtype=torch.float64))
        w = (1 - x) * torch.cos(math.pi * x) + 1.0 / math.pi * torch.sin(math.pi * x)
    elif name == "kaiser":
        w = torch.kaiser_window(M, periodic=False, beta=float(args[0]), dtype=torch.float64)
    elif name == "nuttall":
        a = [0.3635819, 0.4891775, 0.1365995, 0.0106411]
        w = a[0] - a[1] * torch.cos(2 * math.pi * n / (M - 1)) + a[2] * torch.cos(4 * math.pi * n / (M - 1)) - a[3] * torch.cos(6 * math.pi * n / (M - 1))
    elif name == "taylor":
        # Taylor window (nbar nearly-constant sidelobes at `sll` dB below the main lobe), normalised to 1 at the centre
        nbar = int(args[0]) if len(args) > 0 else 4
        sll = float(args[1]) if len(args) > 1 else 30.0
        norm = bool(args[2]) if len(args) > 2 else True
        B_ = 10 ** (sll / 20)
        A_ = math.acosh(B_) / math.pi
        s2 = nbar ** 2 / (A_ ** 2 + (nbar - 0.5) ** 2)
        ma = torch.arange(1, nbar, dtype=torch.float64)
        Fm = torch.zeros(nbar - 1, dtype=torch.float64)
        signs = torch.ones(nbar - 1, dtype=torch.float64)
        signs[1::2] = -1
        m2 = ma * ma
        for mi in range(nbar - 1):
            numer = signs[mi] * torch.prod(1 - m2[mi] / s2 / (A_ ** 2 + (ma - 0.5) ** 2))
            denom = 2 * torch.prod(1 - m2[mi] / m2[:mi]) * torch.prod(1 - m2[mi] / m2[mi + 1:])
            Fm[mi] = numer / denom

        def W(nn_):
            return 1 + 2 * (Fm[None, :] * torch.cos(2 * math.pi * ma[None, :] * (nn_[:, None] - M / 2.0 + 0.5) / M)).sum(1)

        w = W(n)
        if norm:
            w = w / W(torch.tensor([(M - 1) / 2.0], dtype=torch.float64))
    else:
        raise ValueError(f"unknown window {name}")
    if fftbins:
        w = w[:-1]
    from ..framework.dtype import convert_dtype

    return _w(w.to(convert_dtype(dtype)))
