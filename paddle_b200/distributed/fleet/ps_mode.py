"""fleet parameter-server mode. Parity: python/paddle/distributed/fleet/fleet.py (init_server / run_server / init_worker / stop_worker),
base/role_maker.py:PaddleCloudRoleMaker (TRAINING_ROLE / PADDLE_PSERVERS_IP_PORT_LIST / PADDLE_TRAINERS_NUM env protocol),
paddle/fluid/distributed/ps/ (brpc servers, sparse / dense tables).

Servers hold `distributed.ps` tables; transport is `distributed.rpc` (TensorPipe) instead of brpc.  Sparse ids are routed to
`id % n_servers`, dense tables to `hash(name) % n_servers`.  `DistributedStrategy.a_sync = True` makes gradient pushes
fire-and-forget (asynchronous SGD); otherwise a push returns after the server applied it."""
from __future__ import annotations

import os
import zlib

import numpy as np
import torch

from .. import ps as _ps
from .. import rpc as _rpc

_ctx = {"role": None}


class PSRole:
    def __init__(self):
        role = os.environ.get("TRAINING_ROLE", "TRAINER").upper()
        self.endpoints = [e for e in os.environ.get("PADDLE_PSERVERS_IP_PORT_LIST", "").split(",") if e]
        self.n_servers = len(self.endpoints)
        self.n_workers = int(os.environ.get("PADDLE_TRAINERS_NUM", "1"))
        self.is_server = role == "PSERVER"
        if self.is_server:
            me = f"{os.environ.get('POD_IP', '127.0.0.1')}:{os.environ.get('PADDLE_PORT', '')}"
            self.index = self.endpoints.index(me) if me in self.endpoints else int(os.environ.get("PADDLE_PSERVER_ID", "0"))
        else:
            self.index = int(os.environ.get("PADDLE_TRAINER_ID", "0"))

    @property
    def name(self):
        return f"{'server' if self.is_server else 'worker'}{self.index}"

    @property
    def rank(self):
        return self.index if self.is_server else self.n_servers + self.index


def ps_env_present():
    return bool(os.environ.get("PADDLE_PSERVERS_IP_PORT_LIST")) and os.environ.get("TRAINING_ROLE", "").upper() in ("PSERVER", "TRAINER")


def init(strategy=None):
    role = PSRole()
    _ctx.update(role=role, strategy=strategy, server=None, client=None, rpc=False)
    return role


def role():
    return _ctx["role"]


def _ensure_rpc():
    r = _ctx["role"]
    if not _ctx["rpc"]:
        _rpc.init_rpc(r.name, rank=r.rank, world_size=r.n_servers + r.n_workers, master_endpoint=r.endpoints[0])
        _ctx["rpc"] = True


# ---- server side ----------------------------------------------------------------------------------------------------------------
def init_server(dirname=None, var_names=None, **kwargs):
    srv = _ps.ParameterServer()
    _ctx["server"] = srv
    if dirname:
        load_tables(srv, dirname, _ctx["role"].index)
    return srv


def run_server():
    """Serve until every worker called stop_worker (the RPC layer's graceful shutdown is the barrier)."""
    _ensure_rpc()
    _rpc.shutdown()
    _ctx["rpc"] = False


def _srv_create(kind, name, args):
    srv = _ps.ParameterServer._instance
    if name not in srv.tables:
        (srv.create_sparse if kind == "sparse" else srv.create_dense)(name, *args[0], **args[1])
    return True


def _srv_pull_dense(name):
    return _ps.ParameterServer._instance.tables[name].pull()


def _srv_push_dense(name, grad):
    _ps.ParameterServer._instance.tables[name].push(np.asarray(grad))
    return True


def _srv_set_dense(name, value):
    t = _ps.ParameterServer._instance.tables[name]
    with t._lock:
        t.value[...] = np.asarray(value)
    return True


def _srv_table_size(name):
    t = _ps.ParameterServer._instance.tables.get(name)
    return 0 if t is None else (t.size() if hasattr(t, "size") else int(t.value.size))


def _srv_save(dirname, index):
    save_tables(_ps.ParameterServer._instance, dirname, index)
    return True


def save_tables(srv, dirname, index):
    os.makedirs(dirname, exist_ok=True)
    blob = {}
    for name, t in srv.tables.items():
        if isinstance(t, _ps.SparseTable):
            ids = np.array(sorted(t.rows), dtype=np.int64)
            blob[f"sparse::{name}::ids"] = ids
            blob[f"sparse::{name}::rows"] = np.stack([t.rows[int(i)] for i in ids]) if len(ids) else np.zeros((0, t.dim), np.float32)
        else:
            blob[f"dense::{name}"] = t.value
    np.savez(os.path.join(dirname, f"ps_tables_{index}.npz"), **blob)


def load_tables(srv, dirname, index):
    path = os.path.join(dirname, f"ps_tables_{index}.npz")
    if not os.path.exists(path):
        return
    data = np.load(path)
    for key in data.files:
        parts = key.split("::")
        if parts[0] == "dense":
            srv.create_dense(parts[1], data[key].shape)
            srv.tables[parts[1]].value[...] = data[key]
        elif parts[2] == "ids":
            rows = data[f"sparse::{parts[1]}::rows"]
            srv.create_sparse(parts[1], rows.shape[1] if rows.ndim == 2 else 1)
            for i, r in zip(data[key], rows):
                srv.tables[parts[1]].rows[int(i)] = r.astype(np.float32).copy()


# ---- worker side ----------------------------------------------------------------------------------------------------------------
class PSClient:
    """Routes table traffic of one trainer to the servers."""

    def __init__(self, role, a_sync=False):
        self.role, self.a_sync = role, a_sync
        self._pending = []

    def _server_of_dense(self, name):
        return f"server{zlib.crc32(name.encode()) % self.role.n_servers}"

    def create_sparse_table(self, name, dim, **kw):
        for s in range(self.role.n_servers):
            _rpc.rpc_sync(f"server{s}", _srv_create, args=("sparse", name, ((dim,), kw)))

    def create_dense_table(self, name, shape, init=None, **kw):
        _rpc.rpc_sync(self._server_of_dense(name), _srv_create, args=("dense", name, ((tuple(shape),), kw)))
        if init is not None and self.role.index == 0:
            _rpc.rpc_sync(self._server_of_dense(name), _srv_set_dense, args=(name, np.asarray(init, np.float32)))

    def pull_sparse(self, name, ids):
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        n = self.role.n_servers
        out = None
        futs = []
        for s in range(n):
            sel = np.nonzero(ids % n == s)[0]
            if len(sel):
                futs.append((sel, _rpc.rpc_async(f"server{s}", _ps.ParameterServer._pull_sparse, args=(name, ids[sel].tolist()))))
        for sel, f in futs:
            rows = np.asarray(f.wait())
            if out is None:
                out = np.zeros((len(ids), rows.shape[1]), np.float32)
            out[sel] = rows
        return out

    def push_sparse(self, name, ids, grads):
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        grads = np.asarray(grads, np.float32)
        # merge duplicate ids before they leave the trainer
        uniq, inv = np.unique(ids, return_inverse=True)
        merged = np.zeros((len(uniq), grads.shape[1]), np.float32)
        np.add.at(merged, inv, grads)
        n = self.role.n_servers
        for s in range(n):
            sel = np.nonzero(uniq % n == s)[0]
            if len(sel):
                f = _rpc.rpc_async(f"server{s}", _ps.ParameterServer._push_sparse, args=(name, uniq[sel].tolist(), merged[sel]))
                self._pending.append(f) if self.a_sync else f.wait()

    def pull_dense(self, name):
        return np.asarray(_rpc.rpc_sync(self._server_of_dense(name), _srv_pull_dense, args=(name,)))

    def push_dense(self, name, grad):
        f = _rpc.rpc_async(self._server_of_dense(name), _srv_push_dense, args=(name, np.asarray(grad, np.float32)))
        self._pending.append(f) if self.a_sync else f.wait()

    def flush(self):
        for f in self._pending:
            f.wait()
        self._pending = []

    def table_size(self, name):
        return sum(int(_rpc.rpc_sync(f"server{s}", _srv_table_size, args=(name,))) for s in range(self.role.n_servers))

    def save(self, dirname):
        self.flush()
        for s in range(self.role.n_servers):
            _rpc.rpc_sync(f"server{s}", _srv_save, args=(dirname, s))


def init_worker():
    _ensure_rpc()
    st = _ctx.get("strategy")
    _ctx["client"] = PSClient(_ctx["role"], a_sync=bool(getattr(st, "a_sync", False)))
    return _ctx["client"]


def client():
    if _ctx.get("client") is None:
        init_worker()
    return _ctx["client"]


def stop_worker():
    c = _ctx.get("client")
    if c is not None:
        c.flush()
    if _ctx.get("rpc"):
        _rpc.shutdown()
        _ctx["rpc"] = False


# ---- layers whose parameters live on the servers ------------------------------------------------------------------------------------
class _SparsePull(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ids, name):
        c = client()
        flat = ids.reshape(-1).cpu().numpy()
        rows = torch.from_numpy(c.pull_sparse(name, flat)).to(anchor.device)
        ctx.ids, ctx.name = flat, name
        return rows.reshape(*ids.shape, rows.shape[-1])

    @staticmethod
    def backward(ctx, g):
        client().push_sparse(ctx.name, ctx.ids, g.reshape(len(ctx.ids), -1).float().cpu().numpy())
        return None, None, None


from ...nn.layer import Layer as _Layer  # noqa: E402


class DistributedEmbedding(_Layer):
    """Embedding whose rows live in a server-side sparse table: forward pulls the rows of the batch, backward pushes their gradients
    (the server applies its own optimizer). Parity: static.nn.sparse_embedding / fleet's distributed lookup table."""

    def __init__(self, name, dim, **table_kw):
        super().__init__()
        self.name, self.dim = name, dim
        client().create_sparse_table(name, dim, **table_kw)
        self._anchor = self.create_parameter([1], is_bias=True)     # gives the pulled rows a grad path; never updated itself

    def forward(self, ids):
        from ...tensor import Tensor

        raw = ids.as_subclass(torch.Tensor) if isinstance(ids, torch.Tensor) else torch.as_tensor(ids)
        return _SparsePull.apply(self._anchor, raw, self.name).as_subclass(Tensor)
