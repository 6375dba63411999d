"""OpTest-style harness: numpy forward reference + numeric (central-difference) gradient check of an op.
Parity: test/legacy_test/op_test.py (check_output / check_grad with get_numeric_gradient)."""
import numpy as np

import paddle_b200 as paddle


def numeric_grad(fn, inputs, wrt, delta=5e-3):
    """d sum(fn(inputs)) / d inputs[wrt] by central differences (float64)."""
    base = [np.array(a, dtype=np.float64) for a in inputs]
    g = np.zeros_like(base[wrt])
    it = np.nditer(base[wrt], flags=["multi_index"])
    while not it.finished:
        i = it.multi_index
        old = base[wrt][i]
        base[wrt][i] = old + delta
        hi = float(np.sum(fn(*base)))
        base[wrt][i] = old - delta
        lo = float(np.sum(fn(*base)))
        base[wrt][i] = old
        g[i] = (hi - lo) / (2 * delta)
        it.iternext()
    return g


class OpTest:
    """Subclass sets `op` (callable on paddle tensors), `ref` (callable on numpy arrays) and `inputs` (list of float arrays)."""

    op = None
    ref = None
    inputs = ()
    rtol, atol, grad_rtol = 1e-5, 1e-6, 5e-3

    def check_output(self):
        out = type(self).op(*[paddle.to_tensor(np.asarray(a)) for a in self.inputs])
        np.testing.assert_allclose(out.numpy(), type(self).ref(*[np.asarray(a) for a in self.inputs]), rtol=self.rtol, atol=self.atol)

    def check_grad(self, wrt=None):
        wrt = range(len(self.inputs)) if wrt is None else wrt
        ts = [paddle.to_tensor(np.asarray(a, dtype=np.float64), stop_gradient=False) for a in self.inputs]
        type(self).op(*ts).sum().backward()
        for i in wrt:
            num = numeric_grad(lambda *a: type(self).ref(*a), self.inputs, i)
            np.testing.assert_allclose(ts[i].grad.numpy(), num, rtol=self.grad_rtol, atol=1e-4)
