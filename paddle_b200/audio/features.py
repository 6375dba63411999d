"""Parity: python/paddle/audio/features/layers.py (Spectrogram, MelSpectrogram, LogMelSpectrogram, MFCC)."""
from __future__ import annotations

import torch

from ..nn.layer import Layer
from ..tensor import Tensor
from . import functional as AF


class Spectrogram(Layer):
    def __init__(self, n_fft=512, hop_length=None, win_length=None, window="hann", power=1.0, center=True, pad_mode="reflect", dtype="float32"):
        super().__init__()
        self.n_fft, self.hop = n_fft, hop_length or n_fft // 4
        self.win_length, self.power, self.center, self.pad_mode = win_length or n_fft, power, center, pad_mode
        self.register_buffer("fft_window", AF.get_window(window, self.win_length, fftbins=True, dtype=dtype), persistable=False)

    def forward(self, x):
        xr = x.as_subclass(torch.Tensor)
        s = torch.stft(xr, self.n_fft, self.hop, self.win_length, self.fft_window.as_subclass(torch.Tensor).to(xr.dtype), self.center, self.pad_mode, return_complex=True)
        return (s.abs() ** self.power).as_subclass(Tensor)


class MelSpectrogram(Layer):
    def __init__(self, sr=22050, n_fft=2048, hop_length=512, win_length=None, window="hann", power=2.0, center=True, pad_mode="reflect",
                 n_mels=64, f_min=50.0, f_max=None, htk=False, norm="slaney", dtype="float32"):
        super().__init__()
        self._spec = Spectrogram(n_fft, hop_length, win_length, window, power, center, pad_mode, dtype)
        self.register_buffer("fbank_matrix", AF.compute_fbank_matrix(sr, n_fft, n_mels, f_min, f_max, htk, norm, dtype), persistable=False)

    def forward(self, x):
        return torch.matmul(self.fbank_matrix, self._spec(x))


class LogMelSpectrogram(Layer):
    def __init__(self, sr=22050, n_fft=512, hop_length=None, win_length=None, window="hann", power=2.0, center=True, pad_mode="reflect",
                 n_mels=64, f_min=50.0, f_max=None, htk=False, norm="slaney", ref_value=1.0, amin=1e-10, top_db=None, dtype="float32"):
        super().__init__()
        self._mel = MelSpectrogram(sr, n_fft, hop_length, win_length, window, power, center, pad_mode, n_mels, f_min, f_max, htk, norm, dtype)
        self.ref_value, self.amin, self.top_db = ref_value, amin, top_db

    def forward(self, x):
        return AF.power_to_db(self._mel(x), self.ref_value, self.amin, self.top_db)


class MFCC(Layer):
    def __init__(self, sr=22050, n_mfcc=40, n_fft=512, hop_length=None, win_length=None, window="hann", power=2.0, center=True, pad_mode="reflect",
                 n_mels=64, f_min=50.0, f_max=None, htk=False, norm="slaney", ref_value=1.0, amin=1e-10, top_db=None, dtype="float32"):
        super().__init__()
        self._log_mel = LogMelSpectrogram(sr, n_fft, hop_length, win_length, window, power, center, pad_mode, n_mels, f_min, f_max, htk, norm, ref_value, amin, top_db, dtype)
        self.register_buffer("dct_matrix", AF.create_dct(n_mfcc, n_mels, dtype=dtype), persistable=False)

    def forward(self, x):
        lm = self._log_mel(x)
        return torch.matmul(lm.transpose([0, 2, 1]) if lm.dim() == 3 else lm.t(), self.dct_matrix).transpose([0, 2, 1] if lm.dim() == 3 else [1, 0])
