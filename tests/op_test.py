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


# ---------------------------------------------------------------------------------------------------------------------------------
# dtype x place matrix (reference: OpTest.check_output_with_place / check_grad_with_place run every op on every available place and
# in fp32 / fp64 / fp16 / bf16 with per-dtype tolerances; bf16 results are compared after the uint16 round trip).
# ---------------------------------------------------------------------------------------------------------------------------------
DTYPE_TOL = {"float64": (1e-10, 1e-12), "float32": (1e-5, 1e-6), "float16": (2e-3, 1e-3), "bfloat16": (2e-2, 1e-2)}


def available_places():
    import torch

    places = ["cpu"]
    if torch.cuda.is_available():
        places.append("gpu:0")
    return places


class OpTestMatrix(OpTest):
    """check_output / check_grad over `dtypes` x available places.  Low-precision runs feed inputs rounded to that dtype to the numpy
    reference too (so only the op's own arithmetic is compared), and gradients are checked against the float64 numeric gradient."""

    dtypes = ("float64", "float32", "bfloat16", "float16")
    grad_dtypes = ("float64", "float32")

    @staticmethod
    def _round(a, dtype):
        import torch

        t = torch.as_tensor(np.asarray(a, dtype=np.float64))
        return t.to(getattr(torch, dtype)).to(torch.float64).numpy()

    def check_output_with_place(self, place, dtype):
        rtol, atol = DTYPE_TOL[dtype]
        prev = paddle.get_device()
        paddle.set_device(place)
        try:
            rounded = [self._round(a, dtype) for a in self.inputs]
            ts = [paddle.to_tensor(r).astype(dtype) for r in rounded]
            out = type(self).op(*ts)
            assert out.dtype.is_floating_point          # (a case may promote through its own float64 constants)
            got = out.astype("float64").numpy()
            np.testing.assert_allclose(got, type(self).ref(*rounded), rtol=rtol, atol=atol, err_msg=f"{place} {dtype}")
        finally:
            paddle.set_device(prev)

    def check_grad_with_place(self, place, dtype, wrt=None):
        prev = paddle.get_device()
        paddle.set_device(place)
        try:
            wrt = range(len(self.inputs)) if wrt is None else wrt
            ts = [paddle.to_tensor(np.asarray(a, dtype=np.float64)).astype(dtype) for a in self.inputs]
            for t in ts:
                t.stop_gradient = False
            type(self).op(*ts).sum().backward()
            tol = 5e-3 if dtype == "float64" else 2e-2
            for i in wrt:
                num = numeric_grad(lambda *a: type(self).ref(*a), self.inputs, i)
                np.testing.assert_allclose(ts[i].grad.astype("float64").numpy(), num, rtol=tol, atol=1e-3, err_msg=f"{place} {dtype} d/dx{i}")
        finally:
            paddle.set_device(prev)

    def check_all(self, grad=True):
        ran = []
        for place in available_places():
            for dt in self.dtypes:
                if place == "cpu" and dt == "float16" and getattr(self, "skip_cpu_fp16", False):
                    continue
                self.check_output_with_place(place, dt)
                ran.append((place, dt))
            if grad:
                for dt in self.grad_dtypes:
                    self.check_grad_with_place(place, dt)
        return ran
