"""LoD sequence ops (static.nn.sequence_*) against per-sequence numpy loops. Parity: test/sequence/test_sequence_*.py."""
import numpy as np
import pytest

import paddle_b200 as paddle
from paddle_b200 import base
from paddle_b200.static import nn as snn

rng = np.random.RandomState(5)
LENS = [3, 1, 4]
X = rng.randn(8, 2).astype("float32")


def lodt(data=X, lens=LENS):
    return base.create_lod_tensor(data, [lens])


def seqs(a, lens=LENS):
    out, o = [], 0
    for n in lens:
        out.append(a[o:o + n])
        o += n
    return out


def test_lod_plumbing():
    t = lodt()
    assert t.lod() == [[0, 3, 4, 8]] and t.recursive_sequence_lengths() == [LENS] and t.has_valid_recursive_sequence_lengths()
    t2 = paddle.to_tensor(X).set_lod([[0, 2, 8]])
    assert t2.recursive_sequence_lengths() == [[2, 6]]
    t3 = base.create_lod_tensor([[1, 2], [3]], None)
    assert t3.shape == [3, 1] and t3.lod() == [[0, 2, 3]]
    with pytest.raises(ValueError):
        snn.sequence_pool(paddle.to_tensor(X), "sum")


@pytest.mark.parametrize("kind,fn", [("sum", lambda s: s.sum(0)), ("average", lambda s: s.mean(0)), ("sqrt", lambda s: s.sum(0) / np.sqrt(len(s))),
                                     ("max", lambda s: s.max(0)), ("first", lambda s: s[0]), ("last", lambda s: s[-1])])
def test_sequence_pool(kind, fn):
    np.testing.assert_allclose(snn.sequence_pool(lodt(), kind).numpy(), np.stack([fn(s) for s in seqs(X)]), rtol=1e-6)


def test_first_last_softmax_reverse():
    np.testing.assert_allclose(snn.sequence_first_step(lodt()).numpy(), np.stack([s[0] for s in seqs(X)]))
    np.testing.assert_allclose(snn.sequence_last_step(lodt()).numpy(), np.stack([s[-1] for s in seqs(X)]))
    v = X[:, :1]
    sm = snn.sequence_softmax(lodt(v))
    ref = np.concatenate([np.exp(s - s.max()) / np.exp(s - s.max()).sum() for s in seqs(v)])
    np.testing.assert_allclose(sm.numpy(), ref, rtol=1e-6)
    assert sm.lod() == [[0, 3, 4, 8]]
    np.testing.assert_allclose(snn.sequence_reverse(lodt()).numpy(), np.concatenate([s[::-1] for s in seqs(X)]))


def test_concat_slice_expand():
    Y = rng.randn(6, 2).astype("float32")
    ylens = [1, 2, 3]
    out = snn.sequence_concat([lodt(), lodt(Y, ylens)])
    ref = np.concatenate([np.concatenate([a, b]) for a, b in zip(seqs(X), seqs(Y, ylens))])
    np.testing.assert_allclose(out.numpy(), ref)
    assert out.recursive_sequence_lengths() == [[4, 3, 7]]
    sl = snn.sequence_slice(lodt(), paddle.to_tensor(np.array([[1], [0], [2]])), paddle.to_tensor(np.array([[2], [1], [1]])))
    np.testing.assert_allclose(sl.numpy(), np.concatenate([seqs(X)[0][1:3], seqs(X)[1][0:1], seqs(X)[2][2:3]]))
    assert sl.lod() == [[0, 2, 3, 4]]
    rows = rng.randn(3, 2).astype("float32")
    ex = snn.sequence_expand_as(paddle.to_tensor(rows), lodt())
    np.testing.assert_allclose(ex.numpy(), np.repeat(rows, LENS, 0))
    ex2 = snn.sequence_expand(paddle.to_tensor(rows), lodt())
    np.testing.assert_allclose(ex2.numpy(), np.repeat(rows, LENS, 0))
    xs = base.create_lod_tensor(rng.randn(4, 2).astype("float32"), [[1, 2, 1]])
    ex3 = snn.sequence_expand(xs, base.create_lod_tensor(np.zeros((5, 1), "float32"), [[2, 1, 2]]))
    parts = seqs(xs.numpy(), [1, 2, 1])
    np.testing.assert_allclose(ex3.numpy(), np.concatenate([parts[0], parts[0], parts[1], parts[2], parts[2]]))
    assert ex3.recursive_sequence_lengths() == [[1, 1, 2, 1, 1]]


def test_pad_unpad_reshape_scatter_enumerate():
    padded, lens = snn.sequence_pad(lodt(), paddle.to_tensor(np.array([0.0], "float32")))
    assert padded.shape == [3, 4, 2] and lens.numpy().tolist() == LENS
    for i, s in enumerate(seqs(X)):
        np.testing.assert_allclose(padded.numpy()[i, :len(s)], s)
        assert (padded.numpy()[i, len(s):] == 0).all()
    back = snn.sequence_unpad(padded, lens)
    np.testing.assert_allclose(back.numpy(), X)
    assert back.lod() == [[0, 3, 4, 8]]
    padded5, _ = snn.sequence_pad(lodt(), paddle.to_tensor(np.array([9.0], "float32")), maxlen=5)
    assert padded5.shape == [3, 5, 2] and padded5.numpy()[1, 1, 0] == 9
    r = snn.sequence_reshape(base.create_lod_tensor(np.arange(24, dtype="float32").reshape(6, 4), [[2, 4]]), 8)
    assert r.shape == [3, 8] and r.lod() == [[0, 1, 3]]
    base_t = np.ones((3, 6), "float32")
    idx = base.create_lod_tensor(np.array([[1], [2], [0], [5], [5], [3], [0], [1]]), [LENS])
    upd = base.create_lod_tensor(np.arange(8, dtype="float32").reshape(8, 1), [LENS])
    sc = snn.sequence_scatter(paddle.to_tensor(base_t), idx, upd).numpy()
    ref = base_t.copy()
    o = 0
    for i, n in enumerate(LENS):
        for k in range(o, o + n):
            ref[i, idx.numpy()[k, 0]] += upd.numpy()[k, 0]
        o += n
    np.testing.assert_allclose(sc, ref)
    en = snn.sequence_enumerate(base.create_lod_tensor(np.arange(1, 9).reshape(8, 1), [LENS]), 2, pad_value=0)
    assert en.numpy().tolist() == [[1, 2], [2, 3], [3, 0], [4, 0], [5, 6], [6, 7], [7, 8], [8, 0]]


def test_sequence_conv_matches_per_sequence_window():
    paddle.seed(0)
    out = snn.sequence_conv(lodt(), num_filters=3, filter_size=3, bias_attr=False)
    assert out.shape == [8, 3] and out.lod() == [[0, 3, 4, 8]]
    # the first row of a sequence must not see the previous sequence: compare with the same sequences in another order
    paddle.seed(0)
    perm = np.concatenate([seqs(X)[2], seqs(X)[0], seqs(X)[1]])
    out2 = snn.sequence_conv(base.create_lod_tensor(perm, [[4, 3, 1]]), num_filters=3, filter_size=3, bias_attr=False)
    np.testing.assert_allclose(out2.numpy()[:4], out.numpy()[4:], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out2.numpy()[4:7], out.numpy()[:3], rtol=1e-5, atol=1e-6)
