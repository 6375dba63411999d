"""Model-parallel RNG state tracker. Parity: python/paddle/distributed/fleet/layers/mpu/random.py."""
from __future__ import annotations

import contextlib

import torch

MODEL_PARALLEL_RNG = "model_parallel_rng"


class RNGStatesTracker:
    def __init__(self):
        self.states_ = {}
        self.seeds_ = set()

    def reset(self):
        self.states_, self.seeds_ = {}, set()

    def add(self, name, seed):
        if seed in self.seeds_:
            raise ValueError(f"seed {seed} already exists")
        if name in self.states_:
            raise ValueError(f"state {name} already exists")
        self.seeds_.add(seed)
        cpu = torch.get_rng_state()
        cuda = torch.cuda.get_rng_state() if torch.cuda.is_available() else None
        torch.manual_seed(seed)
        self.states_[name] = (torch.get_rng_state(), torch.cuda.get_rng_state() if cuda is not None else None)
        torch.set_rng_state(cpu)
        if cuda is not None:
            torch.cuda.set_rng_state(cuda)

    def get_states_tracker(self):
        return dict(self.states_)

    def set_states_tracker(self, states):
        self.states_ = dict(states)

    @contextlib.contextmanager
    def rng_state(self, name=MODEL_PARALLEL_RNG):
        if name not in self.states_:
            yield  # tracker not seeded (single-card run): use the global generator
            return
        cpu = torch.get_rng_state()
        cuda = torch.cuda.get_rng_state() if torch.cuda.is_available() else None
        s_cpu, s_cuda = self.states_[name]
        torch.set_rng_state(s_cpu)
        if s_cuda is not None:
            torch.cuda.set_rng_state(s_cuda)
        try:
            yield
        finally:
            self.states_[name] = (torch.get_rng_state(), torch.cuda.get_rng_state() if cuda is not None else None)
            torch.set_rng_state(cpu)
            if cuda is not None:
                torch.cuda.set_rng_state(cuda)


_tracker = RNGStatesTracker()


def get_rng_state_tracker():
    return _tracker


def model_parallel_random_seed(seed=None):
    from . import topology as topo

    hcg = topo.get_hybrid_communicate_group()
    rank = hcg.get_model_parallel_rank() if hcg is not None else 0
    base = seed if seed is not None else 1024
    local_seed = base + 1 + rank * 100
    global_seed = base
    _tracker.reset()
    _tracker.add(MODEL_PARALLEL_RNG, local_seed)
    torch.manual_seed(global_seed)


def determinate_seed(rng_name):
    return hash(rng_name) % (2 ** 31)


def dropout(x, p=0.5, axis=None, rng_name=None, training=True, mode="upscale_in_train", name=None):
    from ...nn import functional as F

    if rng_name is None:
        return F.dropout(x, p, axis, training, mode)
    with _tracker.rng_state(rng_name):
        return F.dropout(x, p, axis, training, mode)
