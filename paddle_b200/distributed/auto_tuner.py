"""Parallel-strategy auto tuner. Parity: python/paddle/distributed/auto_tuner/{tuner,search,prune,cost_model}.py.
Enumerates (dp, mp, pp, sharding, micro-batch, recompute) for a transformer config and ranks them with the analytic cost
model (paddle_b200.cost_model) under the 180 GB HBM budget; `measure=` plugs in real step timings."""
from __future__ import annotations

import itertools

from ..cost_model import CostModel


def _divisors(n):
    return [d for d in range(1, n + 1) if n % d == 0]


def search(num_gpus, hidden, layers, ffn, vocab, seq, global_batch, heads=None, hbm_gb=180.0, bytes_per_param=12, measure=None, top_k=5):
    cm = CostModel()
    n_params = layers * (4 * hidden * hidden + 3 * hidden * ffn) + 2 * vocab * hidden
    out = []
    for mp, pp in itertools.product(_divisors(num_gpus), repeat=2):
        if num_gpus % (mp * pp) or (heads and heads % mp) or layers % pp:
            continue
        rest = num_gpus // (mp * pp)
        for sharding in _divisors(rest):
            dp = rest // sharding
            rep = dp * sharding
            if global_batch % rep:
                continue
            for mbs in (1, 2, 4):
                per_rep = global_batch // rep
                if per_rep % mbs:
                    continue
                acc = per_rep // mbs
                for rc in (False, True):
                    p_local = n_params / (mp * pp)
                    state = p_local * (4 + (bytes_per_param - 4) / sharding)
                    act_layer = mbs * seq * (34 * hidden + 5 * (heads or 1) * 0) * 2 / mp
                    act = (layers / pp) * (mbs * seq * hidden * 2 if rc else act_layer) * (min(acc, pp) if pp > 1 else 1)
                    mem = (state + act) / 2 ** 30
                    if mem > hbm_gb * 0.92:
                        continue
                    tok = mbs * seq
                    flops_layer = 2 * tok * (4 * hidden * hidden + 3 * hidden * ffn) / mp + 4 * tok * seq * hidden / mp
                    t_layer = 3 * flops_layer / (cm.peaks.get("bf16_tflops_sustained", 1400.0) * 1e9 * 0.8)
                    if rc:
                        t_layer *= 4 / 3
                    t_comm = 4 * cm.allreduce_ms(tok * hidden * 2, mp) if mp > 1 else 0.0
                    t_mb = (layers / pp) * (t_layer + t_comm)
                    bubble = (pp - 1) / (acc + pp - 1) if pp > 1 else 0.0
                    t_step = acc * t_mb / (1 - bubble) + cm.allreduce_ms(p_local * 2, rep) + cm.mem_ms(p_local * 16)
                    cfg = dict(dp=dp, mp=mp, pp=pp, sharding=sharding, micro_batch=mbs, accumulate=acc, recompute=rc, mem_gb=round(mem, 1),
                               est_ms=round(t_step, 2), est_tokens_per_s=round(global_batch * seq / t_step * 1e3, 1))
                    if measure is not None:
                        cfg["measured_ms"] = measure(cfg)
                    out.append(cfg)
    key = (lambda c: c.get("measured_ms") or c["est_ms"])
    return sorted(out, key=key)[:top_k]


class AutoTuner:
    def __init__(self, tuner_cfg):
        self.cfg = tuner_cfg
        self.history = []

    def search_once(self):
        res = search(**{k: self.cfg[k] for k in ("num_gpus", "hidden", "layers", "ffn", "vocab", "seq", "global_batch") if k in self.cfg}, heads=self.cfg.get("heads"))
        res = [r for r in res if r not in self.history]
        if not res:
            return None
        self.history.append(res[0])
        return res[0]
