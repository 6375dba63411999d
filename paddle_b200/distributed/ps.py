"""Parameter-server mode (in-process tables over the RPC layer). Parity (role): paddle/fluid/distributed/ps/ (brpc sparse /
dense tables, pull/push), python/paddle/distributed/ps/the_one_ps.py.  Sparse tables are hash maps id -> row with SGD /
Adagrad accessors; workers pull rows for a batch of ids and push gradients; dense tables are plain tensors."""
from __future__ import annotations

import threading

import numpy as np


class SparseTable:
    def __init__(self, dim, initializer="uniform", init_range=0.05, optimizer="sgd", lr=0.05):
        self.dim, self.init_range, self.optimizer, self.lr = dim, init_range, optimizer, lr
        self.rows, self.g2 = {}, {}
        self._lock = threading.Lock()
        self._rng = np.random.RandomState(0)

    def pull(self, ids):
        out = np.empty((len(ids), self.dim), dtype=np.float32)
        with self._lock:
            for i, k in enumerate(ids):
                r = self.rows.get(int(k))
                if r is None:
                    r = self.rows[int(k)] = self._rng.uniform(-self.init_range, self.init_range, self.dim).astype(np.float32)
                out[i] = r
        return out

    def push(self, ids, grads):
        with self._lock:
            for k, g in zip(ids, grads):
                k = int(k)
                r = self.rows.setdefault(k, np.zeros(self.dim, np.float32))
                if self.optimizer == "adagrad":
                    s = self.g2.setdefault(k, np.zeros(self.dim, np.float32))
                    s += g * g
                    r -= self.lr * g / (np.sqrt(s) + 1e-6)
                else:
                    r -= self.lr * g

    def size(self):
        return len(self.rows)


class DenseTable:
    def __init__(self, shape, lr=0.05):
        self.value = np.zeros(shape, np.float32)
        self.lr = lr
        self._lock = threading.Lock()

    def pull(self):
        return self.value.copy()

    def push(self, grad):
        with self._lock:
            self.value -= self.lr * grad


class ParameterServer:
    """Holds tables; exposed to workers either in-process or through paddle_b200.distributed.rpc."""

    _instance = None

    def __init__(self):
        self.tables = {}
        ParameterServer._instance = self

    def create_sparse(self, name, dim, **kw):
        self.tables[name] = SparseTable(dim, **kw)

    def create_dense(self, name, shape, **kw):
        self.tables[name] = DenseTable(shape, **kw)

    @staticmethod
    def _pull_sparse(name, ids):
        return ParameterServer._instance.tables[name].pull(ids)

    @staticmethod
    def _push_sparse(name, ids, grads):
        ParameterServer._instance.tables[name].push(ids, grads)
        return True


class Worker:
    def __init__(self, server="server", local=None):
        self.server, self.local = server, local

    def pull_sparse(self, name, ids):
        if self.local is not None:
            return self.local.tables[name].pull(ids)
        from . import rpc

        return rpc.rpc_sync(self.server, ParameterServer._pull_sparse, args=(name, list(map(int, ids))))

    def push_sparse(self, name, ids, grads):
        if self.local is not None:
            return self.local.tables[name].push(ids, grads)
        from . import rpc

        return rpc.rpc_sync(self.server, ParameterServer._push_sparse, args=(name, list(map(int, ids)), np.asarray(grads)))
