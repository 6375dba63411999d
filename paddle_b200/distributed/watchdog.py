"""Communication watchdog / failure detection. Parity: paddle/phi/core/distributed/comm_task_manager.cc (timeouts, dump of
in-flight collectives), fleet/elastic.  A heartbeat thread publishes `hb/<rank>` to the rendezvous TCPStore; a rank whose
heartbeat is older than `timeout_s`, or a local step that does not advance, raises in every surviving rank instead of
hanging; device-side spins in the peer-memory kernels are bounded as well (csrc/comm/p2p_collectives.cu)."""
from __future__ import annotations

import os
import threading
import time

import torch.distributed as dist


class CommWatchdog:
    def __init__(self, timeout_s=600, interval_s=5, on_failure=None):
        self.timeout_s, self.interval_s, self.on_failure = timeout_s, interval_s, on_failure
        self._stop = threading.Event()
        self._thread = None
        self._step = 0
        self._last_progress = time.time()
        self.log = []          # (seq, name, t_start, t_end) of recent collectives
        self._seq = 0

    def _store(self):
        try:
            return dist.distributed_c10d._get_default_store()
        except Exception:
            return None

    def start(self):
        if self._thread is not None:
            return self
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def tick(self):
        """Call once per training step."""
        self._step += 1
        self._last_progress = time.time()

    def record(self, name):
        self._seq += 1
        self.log.append([self._seq, name, time.time(), None])
        if len(self.log) > 256:
            self.log.pop(0)
        return self._seq

    def done(self, seq):
        for e in reversed(self.log):
            if e[0] == seq:
                e[3] = time.time()
                break

    def _run(self):
        store = self._store()
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        while not self._stop.wait(self.interval_s):
            now = time.time()
            if store is not None:
                try:
                    store.set(f"hb/{rank}", str(now))
                    for r in range(world):
                        if r == rank:
                            continue
                        try:
                            t = float(store.get(f"hb/{r}").decode())
                        except Exception:
                            continue
                        if now - t > self.timeout_s:
                            self._fail(f"rank {r} heartbeat is {now - t:.0f}s old")
                except Exception:
                    pass
            if now - self._last_progress > self.timeout_s:
                pending = [e for e in self.log if e[3] is None]
                self._fail(f"no training progress for {now - self._last_progress:.0f}s; in-flight collectives: {pending[-4:]}")

    def _fail(self, why):
        msg = f"[paddle_b200 watchdog] rank {dist.get_rank() if dist.is_initialized() else 0}: {why}"
        if self.on_failure is not None:
            self.on_failure(msg)
        else:
            print(msg, flush=True)
            os._exit(17)

    def stop(self):
        self._stop.set()


_global = [None]


def enable(timeout_s=600, interval_s=5):
    if _global[0] is None:
        _global[0] = CommWatchdog(timeout_s, interval_s).start()
    return _global[0]
