"""RNG state. Parity: python/paddle/framework/random.py."""
from __future__ import annotations

import random as _pyrandom

import numpy as np
import torch


def seed(seed: int):
    s = int(seed)
    torch.manual_seed(s)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(s)
    np.random.seed(s % (2**32))
    _pyrandom.seed(s)
    return torch.default_generator


def get_rng_state(device=None):
    if device is not None and str(device).startswith(("gpu", "cuda")) and torch.cuda.is_available():
        return [torch.cuda.get_rng_state(i) for i in range(torch.cuda.device_count())]
    return [torch.get_rng_state()]


def set_rng_state(state_list, device=None):
    if device is not None and str(device).startswith(("gpu", "cuda")) and torch.cuda.is_available():
        for i, s in enumerate(state_list):
            torch.cuda.set_rng_state(s, i)
    else:
        torch.set_rng_state(state_list[0])


def get_cuda_rng_state():
    if not torch.cuda.is_available():
        return []
    return [torch.cuda.get_rng_state(i) for i in range(torch.cuda.device_count())]


def set_cuda_rng_state(state_list):
    for i, s in enumerate(state_list):
        torch.cuda.set_rng_state(s, i)
