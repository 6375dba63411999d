"""Static pipeline schedules as per-stage op lists, plus the simulator that builds / checks them.

An op is ``(kind, chunk, micro_batch)`` with kind in {"F", "B", "W"}:
  F  forward of model chunk `chunk` (global stage g = chunk * S + stage) on a micro-batch,
  B  backward; in zero-bubble schedules only the input-gradient half (weight gradients are parked, kernels/wgrad.py),
  W  the parked weight-gradient GEMMs of that (chunk, micro-batch).

Every rank builds the lists of ALL stages from (mode, S, V, M) alone, so send / receive orders agree without negotiation.

Parity: the schedules of /root/reference/python/paddle/distributed/fleet/meta_parallel/pipeline_parallel.py
(PipelineParallel 1F1B :forward_backward_pipeline, FThenB, PipelineParallelWithInterleave :2256 FthenB variant) and
/root/reference/python/paddle/distributed/passes/pipeline_scheduler_pass/pipeline_zero_bubble.py:62 (ZB-H1); here they are data
(op lists) consumed by one engine instead of one hand-written loop per schedule.
"""
from __future__ import annotations


def _fthenb(S, V, M):
    out = []
    for s in range(S):
        ops = [("F", v, m) for v in range(V) for m in range(M)]
        ops += [("B", v, m) for v in range(V - 1, -1, -1) for m in range(M)]
        out.append(ops)
    return out


def _one_f_one_b(S, M):
    out = []
    for s in range(S):
        warm = min(S - s - 1, M)
        ops = [("F", 0, i) for i in range(warm)]
        for i in range(M - warm):
            ops += [("F", 0, warm + i), ("B", 0, i)]
        ops += [("B", 0, M - warm + i) for i in range(warm)]
        out.append(ops)
    return out


def _interleaved(S, V, M):
    """Interleaved 1F1B (virtual pipeline): micro-batches advance in groups of S through the V chunks of a rank."""
    if M % S:
        raise ValueError(f"interleaved schedule needs accumulate_steps ({M}) divisible by the pipeline degree ({S})")
    total = M * V

    def fwd_item(k):
        return (k // S) % V, (k // (S * V)) * S + k % S

    def bwd_item(k):
        v, m = fwd_item(k)
        return V - 1 - v, m

    out = []
    for s in range(S):
        warm = min((S - s - 1) * 2 + (V - 1) * S, total)
        ops = [("F",) + fwd_item(k) for k in range(warm)]
        for i in range(total - warm):
            ops += [("F",) + fwd_item(warm + i), ("B",) + bwd_item(i)]
        ops += [("B",) + bwd_item(total - warm + i) for i in range(warm)]
        out.append(ops)
    return out


def _zero_bubble(S, M, cost=(1.0, 1.0, 1.0), max_pending_w=None, max_inflight=None):
    """ZB-H1 style list scheduling: B (input gradient) is split from W (weight gradient); a stage runs a ready B first, then a ready F
    (bounded number of micro-batches in flight), and fills every remaining gap with a parked W.  `max_pending_w` bounds the number of
    micro-batches whose W is outstanding (their activations and output gradients stay alive until then)."""
    cf, cb, cw = cost
    max_pending_w = max_pending_w if max_pending_w is not None else max(2, S)
    t = [0.0] * S
    f_done = [[None] * M for _ in range(S)]     # finish times
    b_done = [[None] * M for _ in range(S)]
    nf, nb, nw = [0] * S, [0] * S, [0] * S
    ops = [[] for _ in range(S)]
    limit = [(max_inflight if max_inflight is not None else S - s) for s in range(S)]   # micro-batches with F done and B not done
    remaining = 3 * S * M
    while remaining:
        # candidate (start_time, priority, stage, kind) for every stage; run the globally earliest one (keeps the simulation causal)
        best = None
        for s in range(S):
            cands = []
            if nb[s] < nf[s]:
                m = nb[s]
                dep = b_done[s + 1][m] if s + 1 < S else f_done[s][m]
                if dep is not None:
                    cands.append((max(t[s], dep), 0, "B"))
            if nf[s] < M and nf[s] - nb[s] < limit[s] and nf[s] - nw[s] < limit[s] + max_pending_w:
                m = nf[s]
                dep = f_done[s - 1][m] if s > 0 else 0.0
                if dep is not None:
                    cands.append((max(t[s], dep), 1, "F"))
            if nw[s] < nb[s]:
                cands.append((t[s], 2, "W"))
            if not cands:
                continue
            ready_now = [c for c in cands if c[0] <= t[s] + 1e-9]
            if ready_now:
                # B/F that can start right now beat W; W only runs when the stage would otherwise idle, or when too many are parked
                pick = min(ready_now, key=lambda c: c[1])
                if pick[2] != "W" and nb[s] - nw[s] >= max_pending_w and any(c[2] == "W" for c in cands):
                    pick = next(c for c in cands if c[2] == "W")
            else:
                w = [c for c in cands if c[2] == "W"]
                pick = w[0] if w else min(cands, key=lambda c: (c[0], c[1]))
            key = (pick[0], pick[1], s)
            if best is None or key < best[0]:
                best = (key, s, pick)
        assert best is not None, "zero-bubble scheduler deadlocked"
        _, s, (start, _, kind) = best
        if kind == "F":
            m = nf[s]
            t[s] = start + cf
            f_done[s][m] = t[s]
            nf[s] += 1
        elif kind == "B":
            m = nb[s]
            t[s] = start + cb
            b_done[s][m] = t[s]
            nb[s] += 1
        else:
            m = nw[s]
            t[s] = start + cw
            nw[s] += 1
        ops[s].append((kind, 0, m))
        remaining -= 1
    return ops


def build(mode, S, M, V=1, **kw):
    """Op lists for every stage. mode: '1F1B' | 'FThenB' | 'VPP' (interleaved 1F1B) | 'ZBH1'."""
    mode = str(mode).upper().replace("-", "").replace("_", "")
    if mode in ("FTHENB",):
        return _fthenb(S, V, M)
    if mode in ("VPP", "INTERLEAVE", "INTERLEAVED", "1F1BINTERLEAVE") or (mode == "1F1B" and V > 1):
        return _interleaved(S, V, M)
    if mode == "1F1B":
        return _one_f_one_b(S, M)
    if mode in ("ZBH1", "ZEROBUBBLE", "ZB"):
        if V != 1:
            raise ValueError("the zero-bubble schedule is defined for one model chunk per rank")
        return _zero_bubble(S, M, **kw)
    raise ValueError(f"unknown pipeline schedule_mode {mode!r}")


def simulate(ops, S, V=1, cost=(1.0, 1.0, 1.0), merged_w=False):
    """Replays per-stage op lists with asynchronous sends; returns (makespan, busy time per stage). Raises on a deadlock or on an op
    whose producer never runs.  merged_w: B ops include the weight gradient (schedules without W ops)."""
    cf, cb, cw = cost
    G = S * V
    pos = [0] * S
    t = [0.0] * S
    busy = [0.0] * S
    done = {}
    left = sum(len(o) for o in ops)
    while left:
        progressed = False
        for s in range(S):
            while pos[s] < len(ops[s]):
                kind, v, m = ops[s][pos[s]]
                g = v * S + s
                if kind == "F":
                    dep = done.get(("F", g - 1, m)) if g > 0 else 0.0
                    dur = cf
                elif kind == "B":
                    own = done.get(("F", g, m))
                    up = done.get(("B", g + 1, m)) if g < G - 1 else own
                    dep = None if (own is None or up is None) else max(own, up)
                    dur = cb + (cw if merged_w else 0.0)
                else:
                    dep = done.get(("B", g, m))
                    dur = cw
                if dep is None:
                    break
                t[s] = max(t[s], dep) + dur
                busy[s] += dur
                done[(kind, g, m)] = t[s]
                pos[s] += 1
                left -= 1
                progressed = True
        if not progressed:
            raise RuntimeError(f"pipeline schedule deadlocks at positions {pos}")
    return max(t), busy
