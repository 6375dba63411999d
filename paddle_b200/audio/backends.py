"""WAV IO backend. Parity: python/paddle/audio/backends/wave_backend.py."""
from __future__ import annotations

import wave

import numpy as np
import torch

from ..tensor import Tensor


class AudioInfo:
    def __init__(self, sample_rate, num_frames, num_channels, bits_per_sample, encoding):
        self.sample_rate, self.num_frames, self.num_channels = sample_rate, num_frames, num_channels
        self.bits_per_sample, self.encoding = bits_per_sample, encoding


def info(filepath):
    with wave.open(filepath, "rb") as f:
        return AudioInfo(f.getframerate(), f.getnframes(), f.getnchannels(), f.getsampwidth() * 8, "PCM_S")


def load(filepath, frame_offset=0, num_frames=-1, normalize=True, channels_first=True):
    with wave.open(filepath, "rb") as f:
        sr, nch, width = f.getframerate(), f.getnchannels(), f.getsampwidth()
        f.setpos(frame_offset)
        raw = f.readframes(f.getnframes() - frame_offset if num_frames < 0 else num_frames)
    dt = {1: np.uint8, 2: np.int16, 4: np.int32}[width]
    a = np.frombuffer(raw, dtype=dt).reshape(-1, nch)
    if normalize:
        a = (a.astype(np.float32) - (128 if width == 1 else 0)) / float(2 ** (8 * width - 1))
    t = torch.from_numpy(np.ascontiguousarray(a.T if channels_first else a))
    return t.as_subclass(Tensor), sr


def save(filepath, src, sample_rate, channels_first=True, encoding=None, bits_per_sample=16):
    a = src.detach().cpu().as_subclass(torch.Tensor).numpy()
    if channels_first:
        a = a.T
    if a.dtype.kind == "f":
        a = np.clip(a, -1.0, 1.0) * (2 ** (bits_per_sample - 1) - 1)
    a = a.astype({8: np.uint8, 16: np.int16, 32: np.int32}[bits_per_sample])
    with wave.open(filepath, "wb") as f:
        f.setnchannels(a.shape[1] if a.ndim > 1 else 1)
        f.setsampwidth(bits_per_sample // 8)
        f.setframerate(sample_rate)
        f.writeframes(a.tobytes())


def list_available_backends():
    return ["wave_backend"]


def get_current_backend():
    return "wave_backend"


def set_backend(name):
    if name != "wave_backend":
        raise NotImplementedError("only the built-in wave backend is available offline")
