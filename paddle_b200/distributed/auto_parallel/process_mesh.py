"""ProcessMesh. Parity: python/paddle/distributed/auto_parallel/process_mesh.py."""
from __future__ import annotations

import numpy as np

_global_mesh = [None]


class ProcessMesh:
    def __init__(self, mesh, dim_names=None, shape=None, process_ids=None):
        if mesh is None:
            mesh = np.asarray(process_ids).reshape(shape)
        self._mesh = np.asarray(mesh, dtype=np.int64)
        if self._mesh.ndim == 0:
            self._mesh = self._mesh.reshape(1)
        self._dim_names = list(dim_names) if dim_names is not None else [f"d{i}" for i in range(self._mesh.ndim)]
        assert len(self._dim_names) == self._mesh.ndim
        self._groups = {}

    shape = property(lambda self: list(self._mesh.shape))
    ndim = property(lambda self: self._mesh.ndim)
    dim_names = property(lambda self: list(self._dim_names))
    mesh = property(lambda self: self._mesh)
    process_ids = property(lambda self: self._mesh.reshape(-1).tolist())

    def __eq__(self, o):
        return isinstance(o, ProcessMesh) and self._mesh.shape == o._mesh.shape and (self._mesh == o._mesh).all() and self._dim_names == o._dim_names

    def __hash__(self):
        return hash((self._mesh.tobytes(), tuple(self._dim_names)))

    def __repr__(self):
        return f"ProcessMesh(shape={self.shape}, process_ids={self.process_ids}, dim_names={self._dim_names})"

    def dim_index(self, d):
        return self._dim_names.index(d) if isinstance(d, str) else int(d)

    def get_dim_size(self, d):
        return self._mesh.shape[self.dim_index(d)]

    def get_mesh_with_dim(self, dim_name, index=None):
        i = self.dim_index(dim_name)
        order = [i] + [j for j in range(self.ndim) if j != i]
        m = self._mesh.transpose(order)
        names = [self._dim_names[j] for j in order]
        if index is not None:
            return ProcessMesh(m[index], names[1:])
        return ProcessMesh(m, names)

    def __getitem__(self, idx):
        sub = self._mesh[idx]
        if np.isscalar(sub) or sub.ndim == 0:
            return ProcessMesh([int(sub)], ["d0"])
        idx_t = idx if isinstance(idx, tuple) else (idx,)
        names = [n for k, n in enumerate(self._dim_names) if k >= len(idx_t) or isinstance(idx_t[k], slice)]
        return ProcessMesh(sub, names)

    def coord_of(self, rank):
        pos = np.argwhere(self._mesh == rank)
        return None if len(pos) == 0 else tuple(int(x) for x in pos[0])

    def ranks_along(self, dim, rank):
        """Ranks of the 1-D sub-mesh through `rank` along mesh dim `dim`."""
        c = self.coord_of(rank)
        if c is None:
            return None
        sl = list(c)
        sl[self.dim_index(dim)] = slice(None)
        return self._mesh[tuple(sl)].tolist()

    def group_along(self, dim):
        """Communication group (created once, collectively) for this rank along `dim`; None for a single-rank dim."""
        from .. import collective, env

        d = self.dim_index(dim)
        if d in self._groups:
            return self._groups[d]
        me = env.get_rank()
        mine = None
        if self._mesh.shape[d] > 1 and env.get_world_size() > 1:
            moved = np.moveaxis(self._mesh, d, -1).reshape(-1, self._mesh.shape[d])
            for row in moved:   # every rank creates every group, in the same order
                g = collective.new_group(row.tolist())
                if me in row.tolist():
                    mine = g
        self._groups[d] = mine
        return mine

    def __enter__(self):
        self._prev = _global_mesh[0]
        _global_mesh[0] = self
        return self

    def __exit__(self, *a):
        _global_mesh[0] = self._prev


def set_mesh(mesh):
    _global_mesh[0] = mesh


def get_mesh():
    return _global_mesh[0]
