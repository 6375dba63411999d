"""SelectedRows + row-sparse embedding gradients through the optimizers and gradient clipping.
Parity model: test/legacy_test/test_selected_rows.py, test_sgd_op.py (SparseSGD), test_adam_op.py (lazy_mode), test_gradient_clip.py."""
import numpy as np
import pytest
import torch

import paddle_b200 as paddle
from paddle_b200.framework.selected_rows import SelectedRows, as_selected_rows


def test_selected_rows_merge_dense_and_algebra():
    v = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    sr = SelectedRows([5, 1, 5, 7], v, height=10)
    assert sr.is_selected_rows() and sr.shape == [10, 3] and sr.has_duplicates()
    m = sr.merge()
    assert m.rows.tolist() == [1, 5, 7] and not m.has_duplicates()
    assert torch.equal(m.value, torch.stack([v[1], v[0] + v[2], v[3]]))
    d = sr.to_dense()
    assert d.shape == (10, 3) and torch.equal(d[5], v[0] + v[2]) and float(d[0].abs().sum()) == 0
    assert torch.allclose(sr.squared_l2_norm(), (d ** 2).sum())
    other = SelectedRows([1, 2], torch.ones(2, 3), 10)
    s = sr.add(other)
    assert s.rows.tolist() == [1, 2, 5, 7] and torch.equal(s.to_dense(), d + other.to_dense())
    assert torch.equal(sr.scale(2.0).to_dense(), 2 * d)
    back = SelectedRows.from_dense(d)
    assert back.rows.tolist() == [1, 5, 7]
    assert torch.equal(sr.to_sparse_coo().to_dense(), d)
    with pytest.raises(ValueError):
        SelectedRows([1, 2, 3], torch.zeros(2, 3), 10)


def _emb_pair(sparse, vocab=50, dim=8, seed=0):
    paddle.seed(seed)
    e = paddle.nn.Embedding(vocab, dim, sparse=sparse)
    return e


def test_sparse_embedding_produces_row_sparse_gradient():
    e = _emb_pair(True)
    ids = paddle.to_tensor(np.array([[3, 7, 3], [9, 7, 0]]))
    out = e(ids)
    (out * out).sum().backward()
    g = torch.Tensor.grad.__get__(e.weight)
    assert g.layout == torch.sparse_coo
    sr = as_selected_rows(g).merge()
    assert sr.rows.tolist() == [0, 3, 7, 9] and sr.height == 50
    ed = _emb_pair(False)
    outd = ed(ids)
    (outd * outd).sum().backward()
    assert torch.allclose(sr.to_dense(), ed.weight.grad.as_subclass(torch.Tensor))


@pytest.mark.parametrize("opt_name", ["sgd", "adam", "adamw_lazy"])
def test_optimizers_with_row_sparse_gradients(opt_name):
    ids = [np.array([[1, 4, 4]]), np.array([[2, 4, 30]]), np.array([[1, 1, 7]])]

    def run(sparse):
        e = _emb_pair(sparse, seed=3)
        if opt_name == "sgd":
            opt = paddle.optimizer.SGD(0.1, parameters=e.parameters())
        elif opt_name == "adam":
            opt = paddle.optimizer.Adam(0.05, parameters=e.parameters())
        else:
            opt = paddle.optimizer.AdamW(0.05, parameters=e.parameters(), weight_decay=0.01, lazy_mode=True)
        w0 = e.weight.numpy().copy()
        for b in ids:
            loss = (e(paddle.to_tensor(b)) ** 2).sum()
            loss.backward()
            opt.step()
            opt.clear_grad()
        return w0, e.weight.numpy(), opt

    w0, ws, opt_s = run(True)
    _, wd, _ = run(False)
    touched = sorted({int(i) for b in ids for i in b.reshape(-1)})
    untouched = [i for i in range(50) if i not in touched]
    if opt_name == "adamw_lazy":
        # lazy: rows that never appeared are bit-identical to the initial weights (no decay, no moment updates) ...
        np.testing.assert_array_equal(ws[untouched], w0[untouched])
        assert not np.allclose(ws[touched], w0[touched])
        # ... and a row's moments only move in the steps it appears: row 30 appears once, its first moment is (1 - beta1) * g
        m = opt_s._accumulators["moment1"][list(opt_s._accumulators["moment1"])[0]]
        assert float(m[untouched].abs().sum()) == 0.0 and float(m[30].abs().sum()) > 0
    else:
        # sgd / non-lazy adam: same result as the dense gradient
        np.testing.assert_allclose(ws, wd, rtol=1e-5, atol=1e-6)
        if opt_name == "sgd":
            np.testing.assert_array_equal(ws[untouched], w0[untouched])


def test_global_norm_clip_counts_merged_rows():
    e = _emb_pair(True, seed=5)
    lin = paddle.nn.Linear(8, 4)
    clip = paddle.nn.ClipGradByGlobalNorm(0.5)
    params = list(e.parameters()) + list(lin.parameters())
    opt = paddle.optimizer.SGD(1.0, parameters=params, grad_clip=clip)
    w0 = e.weight.numpy().copy()
    loss = (lin(e(paddle.to_tensor(np.array([[2, 2, 5]])))) ** 2).sum() * 100.0
    loss.backward()
    sr = as_selected_rows(torch.Tensor.grad.__get__(e.weight)).merge()
    dense_parts = [p.grad.as_subclass(torch.Tensor) for p in lin.parameters()]
    gn = float(torch.sqrt(sr.squared_l2_norm() + sum((d ** 2).sum() for d in dense_parts)))
    assert gn > 0.5
    opt.step()
    moved = e.weight.numpy() - w0
    expect = -(sr.to_dense() * (0.5 / gn)).numpy()
    np.testing.assert_allclose(moved, expect, rtol=1e-4, atol=1e-6)


def test_value_and_norm_clip_on_duplicate_rows():
    sr = SelectedRows([2, 2, 5], torch.tensor([[3.0, -4.0], [3.0, -4.0], [0.5, 0.5]]), height=8)
    p = paddle.to_tensor(np.zeros((8, 2), "float32"))
    (_, c), = paddle.nn.ClipGradByValue(5.0)([(p, sr)])
    assert c.rows.tolist() == [2, 5] and torch.equal(c.value, torch.tensor([[5.0, -5.0], [0.5, 0.5]]))     # rows are summed first (6, -8), then clipped
    (_, n), = paddle.nn.ClipGradByNorm(1.0)([(p, sr)])
    dense = sr.to_dense()
    assert torch.allclose(n.to_dense(), dense / dense.norm(), atol=1e-6)
