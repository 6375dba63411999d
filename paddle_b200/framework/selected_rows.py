"""SelectedRows: a row-sparse tensor (`rows` index into a dense [height, ...] tensor, `value` holds those rows).

Parity: paddle/phi/core/selected_rows.h + the merge / add functors in paddle/phi/kernels/funcs/selected_rows_functor.cu and the sparse
(lazy) branches of the sgd / adam kernels (paddle/phi/kernels/selected_rows/).  It is what `nn.Embedding(sparse=True)` produces as weight
gradient: the optimizers update only the touched rows (Adam / AdamW with lazy_mode=True also keeps the moments of the other rows frozen),
gradient clipping sees the merged rows."""
from __future__ import annotations

import torch


class SelectedRows:
    __slots__ = ("rows", "value", "height")

    def __init__(self, rows, value, height):
        rows = rows if isinstance(rows, torch.Tensor) else torch.as_tensor(list(rows), dtype=torch.int64)
        self.rows = rows.to(torch.int64).reshape(-1)
        self.value = value
        self.height = int(height)
        if self.value.shape[0] != self.rows.numel():
            raise ValueError(f"SelectedRows: {self.rows.numel()} rows but value has {self.value.shape[0]}")

    # ---- construction -------------------------------------------------------------------------------------------------------------
    @classmethod
    def from_sparse_coo(cls, g):
        """torch sparse COO gradient of an embedding weight ([height, d], sparse in dim 0) -> SelectedRows (rows may repeat)."""
        if g.layout != torch.sparse_coo:
            raise TypeError("from_sparse_coo expects a torch.sparse_coo tensor")
        idx = g._indices()
        if idx.shape[0] != 1:
            g = g.coalesce()
            dense_rows = g.to_dense()
            nz = torch.nonzero(dense_rows.reshape(dense_rows.shape[0], -1).abs().sum(1) > 0).reshape(-1)
            return cls(nz, dense_rows[nz], g.shape[0])
        return cls(idx[0], g._values(), g.shape[0])

    @classmethod
    def from_dense(cls, dense, rows=None):
        if rows is None:
            rows = torch.nonzero(dense.reshape(dense.shape[0], -1).abs().sum(1) > 0).reshape(-1)
        rows = torch.as_tensor(rows, dtype=torch.int64, device=dense.device)
        return cls(rows, dense[rows], dense.shape[0])

    # ---- queries --------------------------------------------------------------------------------------------------------------------
    def is_selected_rows(self):
        return True

    @property
    def shape(self):
        return [self.height] + list(self.value.shape[1:])

    @property
    def dtype(self):
        return self.value.dtype

    @property
    def device(self):
        return self.value.device

    def has_duplicates(self):
        return torch.unique(self.rows).numel() != self.rows.numel()

    # ---- algebra --------------------------------------------------------------------------------------------------------------------
    def merge(self):
        """Sum the values of repeated rows; rows come back sorted (MergeAdd)."""
        if self.rows.numel() == 0:
            return SelectedRows(self.rows, self.value, self.height)
        uniq, inv = torch.unique(self.rows, sorted=True, return_inverse=True)
        if uniq.numel() == self.rows.numel() and bool((self.rows[1:] > self.rows[:-1]).all()):
            return self
        out = torch.zeros((uniq.numel(),) + tuple(self.value.shape[1:]), dtype=self.value.dtype, device=self.value.device)
        out.index_add_(0, inv, self.value)
        return SelectedRows(uniq, out, self.height)

    def to_dense(self):
        out = torch.zeros([self.height] + list(self.value.shape[1:]), dtype=self.value.dtype, device=self.value.device)
        out.index_add_(0, self.rows.to(out.device), self.value)
        return out

    def to_sparse_coo(self):
        m = self.merge()
        return torch.sparse_coo_tensor(m.rows.unsqueeze(0), m.value, [self.height] + list(self.value.shape[1:])).coalesce()

    def scale(self, s):
        return SelectedRows(self.rows, self.value * s, self.height)

    def add(self, other):
        if isinstance(other, SelectedRows):
            if other.height != self.height:
                raise ValueError("SelectedRows.add: heights differ")
            return SelectedRows(torch.cat([self.rows, other.rows]), torch.cat([self.value, other.value]), self.height).merge()
        return self.to_dense() + other

    def squared_l2_norm(self):
        return (self.merge().value.float() ** 2).sum()

    def __repr__(self):
        return f"SelectedRows(height={self.height}, rows={self.rows.tolist() if self.rows.numel() <= 16 else str(self.rows.numel()) + ' rows'}, value{tuple(self.value.shape)})"


def as_selected_rows(grad):
    """Gradient of a `sparse=True` embedding (torch sparse COO) as SelectedRows; dense tensors pass through unchanged."""
    if isinstance(grad, SelectedRows):
        return grad
    if isinstance(grad, torch.Tensor) and grad.layout == torch.sparse_coo:
        return SelectedRows.from_sparse_coo(grad)
    return grad
