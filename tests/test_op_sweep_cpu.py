"""Broad OpTest sweep: forward against numpy for ~150 ops, analytic-vs-numeric gradients where the op is differentiable.
Parity: the per-op files under test/legacy_test/test_*_op.py, collapsed into one table."""
import math

import numpy as np
import pytest
import scipy.special as sps

import paddle_b200 as paddle
from op_test import OpTest

rng = np.random.RandomState(7)
A = rng.randn(3, 4)
A = np.sign(A) * (np.abs(A) + 0.05)   # keep away from the kinks at 0 (numeric gradients)
B = rng.randn(3, 4)
P = rng.rand(3, 4) + 0.5          # positive
U = rng.rand(3, 4) * 0.8 + 0.1    # in (0.1, 0.9)
V = rng.randn(4)
M = rng.randn(4, 4)
SPD = M @ M.T + 4 * np.eye(4)
I = rng.randint(0, 5, size=(3, 4))
J = rng.randint(1, 5, size=(3, 4))
F = paddle.nn.functional
SEP = rng.permutation(12).reshape(3, 4) * 0.37 - 2.0   # well separated values (max/min/sort gradients are well defined)

# name: (op, ref, inputs, check_grad)
UNARY = {
    "abs": (paddle.abs, np.abs, [A], True), "exp": (paddle.exp, np.exp, [A], True), "expm1": (paddle.expm1, np.expm1, [A], True),
    "log": (paddle.log, np.log, [P], True), "log2": (paddle.log2, np.log2, [P], True), "log10": (paddle.log10, np.log10, [P], True),
    "log1p": (paddle.log1p, np.log1p, [P], True), "sqrt": (paddle.sqrt, np.sqrt, [P], True), "rsqrt": (paddle.rsqrt, lambda a: 1 / np.sqrt(a), [P], True),
    "square": (paddle.square, np.square, [A], True), "reciprocal": (paddle.reciprocal, np.reciprocal, [P], True),
    "sin": (paddle.sin, np.sin, [A], True), "cos": (paddle.cos, np.cos, [A], True), "tan": (paddle.tan, np.tan, [U], True),
    "asin": (paddle.asin, np.arcsin, [U], True), "acos": (paddle.acos, np.arccos, [U], True), "atan": (paddle.atan, np.arctan, [A], True),
    "sinh": (paddle.sinh, np.sinh, [A], True), "cosh": (paddle.cosh, np.cosh, [A], True), "asinh": (paddle.asinh, np.arcsinh, [A], True),
    "acosh": (paddle.acosh, np.arccosh, [P + 1], True), "atanh": (paddle.atanh, np.arctanh, [U], True),
    "sigmoid": (F.sigmoid, sps.expit, [A], True), "erf": (paddle.erf, sps.erf, [A], True), "erfinv": (paddle.erfinv, sps.erfinv, [U], True),
    "lgamma": (paddle.lgamma, sps.gammaln, [P], True), "digamma": (paddle.digamma, sps.digamma, [P], True),
    "floor": (paddle.floor, np.floor, [A * 3], False), "ceil": (paddle.ceil, np.ceil, [A * 3], False), "round": (paddle.round, np.round, [A * 3], False),
    "trunc": (paddle.trunc, np.trunc, [A * 3], False), "sign": (paddle.sign, np.sign, [A], False), "neg": (paddle.neg, np.negative, [A], True),
    "frac": (paddle.frac, lambda a: a - np.trunc(a), [A * 3], False), "logit": (paddle.logit, sps.logit, [U], True),
    "i0": (paddle.i0, sps.i0, [A], True), "i0e": (paddle.i0e, sps.i0e, [A], True), "i1": (paddle.i1, sps.i1, [A], True), "i1e": (paddle.i1e, sps.i1e, [A], True),
    "relu": (F.relu, lambda a: np.maximum(a, 0), [A], True), "relu6": (F.relu6, lambda a: np.clip(a, 0, 6), [A * 4], True),
    "leaky_relu": (lambda x: F.leaky_relu(x, 0.1), lambda a: np.where(a > 0, a, 0.1 * a), [A], True),
    "elu": (F.elu, lambda a: np.where(a > 0, a, np.exp(a) - 1), [A], True),
    "selu": (F.selu, lambda a: 1.0507009873554805 * np.where(a > 0, a, 1.6732632423543772 * (np.exp(a) - 1)), [A], True),
    "celu": (lambda x: F.celu(x, 1.5), lambda a: np.maximum(a, 0) + np.minimum(0, 1.5 * (np.exp(a / 1.5) - 1)), [A], True),
    "softplus": (F.softplus, lambda a: np.log1p(np.exp(a)), [A], True), "softsign": (F.softsign, lambda a: a / (1 + np.abs(a)), [A], True),
    "silu": (F.silu, lambda a: a * sps.expit(a), [A], True), "swish": (F.swish, lambda a: a * sps.expit(a), [A], True),
    "mish": (F.mish, lambda a: a * np.tanh(np.log1p(np.exp(a))), [A], True),
    "hardtanh": (F.hardtanh, lambda a: np.clip(a, -1, 1), [A * 2], True),
    "hardsigmoid": (F.hardsigmoid, lambda a: np.clip(a * 0.1666667 + 0.5, 0, 1), [A * 4], True),
    "hardswish": (F.hardswish, lambda a: a * np.clip(a + 3, 0, 6) / 6, [A * 4], True),
    "hardshrink": (F.hardshrink, lambda a: np.where(np.abs(a) > 0.5, a, 0), [A], True),
    "softshrink": (F.softshrink, lambda a: np.where(a > 0.5, a - 0.5, np.where(a < -0.5, a + 0.5, 0)), [A], True),
    "tanhshrink": (F.tanhshrink, lambda a: a - np.tanh(a), [A], True), "log_sigmoid": (F.log_sigmoid, lambda a: np.log(sps.expit(a)), [A], True),
    "thresholded_relu": (lambda x: F.thresholded_relu(x, 0.3), lambda a: np.where(a > 0.3, a, 0), [A], True),
    "log_softmax": (lambda x: F.log_softmax(x, -1) * paddle.to_tensor(np.arange(4.0)), lambda a: (a - sps.logsumexp(a, -1, keepdims=True)) * np.arange(4.0), [A], True),
    "stanh": (lambda x: paddle.stanh(x, 0.67, 1.7159), lambda a: 1.7159 * np.tanh(0.67 * a), [A], True),
    "nan_to_num": (paddle.nan_to_num, np.nan_to_num, [np.array([1.0, np.nan, np.inf, -np.inf])], False),
    "isnan": (paddle.isnan, np.isnan, [np.array([1.0, np.nan, np.inf])], False), "isinf": (paddle.isinf, np.isinf, [np.array([1.0, np.nan, np.inf])], False),
    "isfinite": (paddle.isfinite, np.isfinite, [np.array([1.0, np.nan, np.inf])], False),
    "deg2rad": (paddle.deg2rad, np.deg2rad, [A * 90], True), "rad2deg": (paddle.rad2deg, np.rad2deg, [A], True),
    "angle": (paddle.angle, np.angle, [A], False), "sgn": (paddle.sgn, np.sign, [A], False),
}

BINARY = {
    "add": (paddle.add, np.add, [A, B], True), "subtract": (paddle.subtract, np.subtract, [A, B], True), "multiply": (paddle.multiply, np.multiply, [A, B], True),
    "divide": (paddle.divide, np.divide, [A, P], True), "pow": (paddle.pow, np.power, [P, B], True), "maximum": (paddle.maximum, np.maximum, [A, B], True),
    "minimum": (paddle.minimum, np.minimum, [A, B], True), "fmax": (paddle.fmax, np.fmax, [A, B], True), "fmin": (paddle.fmin, np.fmin, [A, B], True),
    "atan2": (paddle.atan2, np.arctan2, [A, B], True), "hypot": (paddle.hypot, np.hypot, [A, B], True), "copysign": (paddle.copysign, np.copysign, [A, B], False),
    "floor_divide": (paddle.floor_divide, np.floor_divide, [I, J], False), "remainder": (paddle.remainder, np.remainder, [A * 3, P], False),
    "mod_int": (paddle.mod, np.mod, [I, J], False), "gcd": (paddle.gcd, np.gcd, [I, J], False), "lcm": (paddle.lcm, np.lcm, [I, J], False),
    "bitwise_and": (paddle.bitwise_and, np.bitwise_and, [I, J], False), "bitwise_or": (paddle.bitwise_or, np.bitwise_or, [I, J], False),
    "bitwise_xor": (paddle.bitwise_xor, np.bitwise_xor, [I, J], False), "logical_and": (paddle.logical_and, np.logical_and, [I > 2, J > 2], False),
    "logical_or": (paddle.logical_or, np.logical_or, [I > 2, J > 2], False), "logical_xor": (paddle.logical_xor, np.logical_xor, [I > 2, J > 2], False),
    "equal": (paddle.equal, np.equal, [I, J], False), "not_equal": (paddle.not_equal, np.not_equal, [I, J], False), "less_than": (paddle.less_than, np.less, [A, B], False),
    "less_equal": (paddle.less_equal, np.less_equal, [I, J], False), "greater_than": (paddle.greater_than, np.greater, [A, B], False),
    "greater_equal": (paddle.greater_equal, np.greater_equal, [I, J], False), "heaviside": (paddle.heaviside, np.heaviside, [A, B], False),
    "logaddexp": (paddle.logaddexp, np.logaddexp, [A, B], True), "nextafter": (paddle.nextafter, np.nextafter, [A, B], False),
    "ldexp": (paddle.ldexp, lambda a, b: a * 2.0 ** b, [A, I], False), "lerp": (lambda x, y: paddle.lerp(x, y, 0.3), lambda a, b: a + 0.3 * (b - a), [A, B], True),
    "dot": (paddle.dot, np.dot, [V, V * 2], True), "outer": (paddle.outer, np.outer, [V, V + 1], True), "inner": (paddle.inner, np.inner, [A, B], True),
    "cross": (lambda x, y: paddle.cross(x, y, axis=1), lambda a, b: np.cross(a, b, axis=1), [A[:, :3], B[:, :3]], True),
    "kron": (paddle.kron, np.kron, [A[:2, :2], B[:2, :3]], True), "mm": (paddle.mm, np.matmul, [A, B.T], True),
    "bmm": (paddle.bmm, np.matmul, [rng.randn(2, 3, 4), rng.randn(2, 4, 2)], True), "mv": (paddle.mv, np.matmul, [A, V], True),
    "addmm": (lambda i, x, y: paddle.addmm(i, x, y, beta=0.5, alpha=2.0), lambda i, a, b: 0.5 * i + 2.0 * a @ b, [rng.randn(3, 3), A, B.T], True),
    "dist": (lambda x, y: paddle.dist(x, y, 2), lambda a, b: np.linalg.norm(a - b), [A, B], True),
}

REDUCE = {
    "sum": (lambda x: paddle.sum(x, axis=1), lambda a: a.sum(1), [A], True), "mean": (lambda x: paddle.mean(x, axis=0), lambda a: a.mean(0), [A], True),
    "prod": (lambda x: paddle.prod(x, axis=1), lambda a: a.prod(1), [P], True), "max": (lambda x: paddle.max(x, axis=1), lambda a: a.max(1), [SEP], True),
    "min": (lambda x: paddle.min(x, axis=1), lambda a: a.min(1), [SEP], True), "amax": (lambda x: paddle.amax(x, axis=1), lambda a: a.max(1), [SEP], True),
    "amin": (lambda x: paddle.amin(x, axis=1), lambda a: a.min(1), [SEP], True), "std": (lambda x: paddle.std(x, axis=1), lambda a: a.std(1, ddof=1), [A], True),
    "var": (lambda x: paddle.var(x, axis=1), lambda a: a.var(1, ddof=1), [A], True), "median": (lambda x: paddle.median(x, axis=1), lambda a: np.median(a, 1), [rng.randn(3, 5)], False),
    "nansum": (paddle.nansum, np.nansum, [np.where(A > 1, np.nan, A)], False), "nanmean": (paddle.nanmean, np.nanmean, [np.where(A > 1, np.nan, A)], False),
    "all": (lambda x: paddle.all(x, axis=1), lambda a: a.all(1), [I > 0], False), "any": (lambda x: paddle.any(x, axis=1), lambda a: a.any(1), [I > 3], False),
    "argmax": (lambda x: paddle.argmax(x, axis=1), lambda a: a.argmax(1), [A], False), "argmin": (lambda x: paddle.argmin(x, axis=1), lambda a: a.argmin(1), [A], False),
    "cumprod": (lambda x: paddle.cumprod(x, dim=1), lambda a: np.cumprod(a, 1), [P], True),
    "logcumsumexp": (lambda x: paddle.logcumsumexp(x, axis=1), lambda a: np.log(np.cumsum(np.exp(a), 1)), [A], True),
    "cummax": (lambda x: paddle.cummax(x, axis=1)[0], lambda a: np.maximum.accumulate(a, 1), [A], False),
    "cummin": (lambda x: paddle.cummin(x, axis=1)[0], lambda a: np.minimum.accumulate(a, 1), [A], False),
    "norm_fro": (lambda x: paddle.linalg.norm(x), lambda a: np.linalg.norm(a), [A], True), "norm_1": (lambda x: paddle.linalg.norm(x, p=1, axis=1), lambda a: np.abs(a).sum(1), [A], True),
    "trace": (paddle.trace, np.trace, [M], True), "count_nonzero": (paddle.count_nonzero, np.count_nonzero, [I], False),
    "quantile": (lambda x: paddle.quantile(x, 0.3, axis=1), lambda a: np.quantile(a, 0.3, axis=1), [A], False),
    "kthvalue": (lambda x: paddle.kthvalue(x, 2, axis=1)[0], lambda a: np.sort(a, 1)[:, 1], [A], False),
    "mode": (lambda x: paddle.mode(x, axis=1)[0], lambda a: np.array([np.bincount(r).argmax() for r in a]), [I], False),
}

SHAPE = {
    "reshape": (lambda x: paddle.reshape(x, [4, 3]), lambda a: a.reshape(4, 3), [A], True), "transpose": (lambda x: paddle.transpose(x, [1, 0]), lambda a: a.T, [A], True),
    "flatten": (paddle.flatten, lambda a: a.reshape(-1), [A], True), "squeeze": (lambda x: paddle.squeeze(x, 0), lambda a: a[0], [A[None]], True),
    "unsqueeze": (lambda x: paddle.unsqueeze(x, 1), lambda a: a[:, None], [A], True), "flip": (lambda x: paddle.flip(x, [1]), lambda a: a[:, ::-1], [A], True),
    "roll": (lambda x: paddle.roll(x, 1, 1), lambda a: np.roll(a, 1, 1), [A], True), "tile": (lambda x: paddle.tile(x, [2, 1]), lambda a: np.tile(a, (2, 1)), [A], True),
    "expand": (lambda x: paddle.expand(x, [2, 3, 4]), lambda a: np.broadcast_to(a, (2, 3, 4)), [A], True),
    "concat": (lambda x, y: paddle.concat([x, y], axis=1), lambda a, b: np.concatenate([a, b], 1), [A, B], True),
    "stack": (lambda x, y: paddle.stack([x, y], axis=0), lambda a, b: np.stack([a, b]), [A, B], True),
    "tril": (paddle.tril, np.tril, [M], True), "triu": (paddle.triu, np.triu, [M], True), "diag": (paddle.diag, np.diag, [V], True),
    "diagonal": (paddle.diagonal, lambda a: np.diagonal(a), [M], True), "rot90": (paddle.rot90, np.rot90, [A], True),
    "moveaxis": (lambda x: paddle.moveaxis(x, 0, 1), lambda a: np.moveaxis(a, 0, 1), [A], True),
    "clip": (lambda x: paddle.clip(x, -0.5, 0.5), lambda a: np.clip(a, -0.5, 0.5), [A], True),
    "where": (lambda x, y: paddle.where(x > 0, x, y), lambda a, b: np.where(a > 0, a, b), [A, B], True),
    "gather": (lambda x: paddle.gather(x, paddle.to_tensor(np.array([2, 0])), axis=0), lambda a: a[[2, 0]], [A], True),
    "index_select": (lambda x: paddle.index_select(x, paddle.to_tensor(np.array([3, 1])), axis=1), lambda a: a[:, [3, 1]], [A], True),
    "take_along_axis": (lambda x: paddle.take_along_axis(x, paddle.to_tensor(np.argsort(A, 1)), 1), lambda a: np.take_along_axis(a, np.argsort(A, 1), 1), [A], True),
    "sort": (lambda x: paddle.sort(x, axis=1), lambda a: np.sort(a, 1), [SEP], True), "argsort": (lambda x: paddle.argsort(x, axis=1), lambda a: np.argsort(a, 1), [A], False),
    "topk": (lambda x: paddle.topk(x, 2, axis=1)[0], lambda a: -np.sort(-a, 1)[:, :2], [SEP], True),
    "masked_fill": (lambda x: paddle.masked_fill(x, paddle.to_tensor(A > 0), 2.0), lambda a: np.where(A > 0, 2.0, a), [A], True),
    "pad": (lambda x: F.pad(x, [1, 2], value=0.5), lambda a: np.pad(a, ((0, 0), (1, 2)), constant_values=0.5), [A], True),
    "repeat_interleave": (lambda x: paddle.repeat_interleave(x, 2, axis=0), lambda a: np.repeat(a, 2, 0), [A], True),
    "diff": (lambda x: paddle.diff(x, axis=1), lambda a: np.diff(a, axis=1), [A], True), "cast": (lambda x: paddle.cast(x, "int32"), lambda a: a.astype("int32"), [A * 3], False),
    "one_hot": (lambda x: F.one_hot(x, 5), lambda a: np.eye(5)[a], [I], False), "bincount": (paddle.bincount, np.bincount, [I.reshape(-1)], False),
    "searchsorted": (lambda s, v: paddle.searchsorted(s, v), np.searchsorted, [np.sort(V), rng.randn(5)], False),
    "bucketize": (lambda v, s: paddle.bucketize(v, s), lambda v, s: np.searchsorted(s, v), [rng.randn(5), np.sort(V)], False),
    "unique": (lambda x: paddle.unique(x), np.unique, [I.reshape(-1)], False), "nonzero": (lambda x: paddle.nonzero(x), lambda a: np.stack(np.nonzero(a), 1), [I > 2], False),
    "meshgrid": (lambda x, y: paddle.meshgrid(x, y)[0], lambda a, b: np.meshgrid(a, b, indexing="ij")[0], [V, V[:3]], False),
    "tensordot": (lambda x, y: paddle.tensordot(x, y, axes=1), lambda a, b: np.tensordot(a, b, 1), [A, B.T], True),
    "einsum": (lambda x, y: paddle.einsum("ij,kj->ik", x, y), lambda a, b: np.einsum("ij,kj->ik", a, b), [A, B], True),
}

LINALG = {
    "inv": (paddle.linalg.inv, np.linalg.inv, [SPD], True), "det": (paddle.linalg.det, np.linalg.det, [SPD / 4], True),
    "slogdet": (lambda x: paddle.linalg.slogdet(x)[1], lambda a: np.linalg.slogdet(a)[1], [SPD], True),
    "cholesky": (paddle.linalg.cholesky, np.linalg.cholesky, [SPD], False), "solve": (paddle.linalg.solve, np.linalg.solve, [SPD, A.T], True),
    "pinv": (paddle.linalg.pinv, np.linalg.pinv, [A], False), "matrix_power": (lambda x: paddle.linalg.matrix_power(x, 3), lambda a: np.linalg.matrix_power(a, 3), [M / 2], True),
    "eigvalsh": (paddle.linalg.eigvalsh, np.linalg.eigvalsh, [SPD], False), "svdvals": (lambda x: paddle.linalg.svd(x)[1], lambda a: np.linalg.svd(a)[1], [A], False),
    "matrix_rank": (paddle.linalg.matrix_rank, np.linalg.matrix_rank, [SPD], False), "cond": (paddle.linalg.cond, np.linalg.cond, [SPD], False),
    "multi_dot": (lambda x, y, z: paddle.linalg.multi_dot([x, y, z]), lambda a, b, c: a @ b @ c, [A, B.T, A], True),
    "triangular_solve": (lambda x, y: paddle.linalg.triangular_solve(x, y, upper=False), lambda a, b: np.linalg.solve(np.tril(a), b), [np.tril(SPD), A.T], False),
    "cov": (paddle.linalg.cov, np.cov, [A], False), "corrcoef": (paddle.linalg.corrcoef, np.corrcoef, [A], False),
}

ALL = {}
for group, table in (("unary", UNARY), ("binary", BINARY), ("reduce", REDUCE), ("shape", SHAPE), ("linalg", LINALG)):
    for k, v in table.items():
        ALL[f"{group}.{k}"] = v


@pytest.mark.parametrize("name", sorted(ALL))
def test_op(name):
    op, ref, inputs, grad = ALL[name]

    class T(OpTest):
        pass

    T.op, T.ref, T.inputs = staticmethod(op), staticmethod(ref), inputs
    t = T()
    t.rtol, t.atol = 1e-6, 1e-8
    t.check_output()
    if grad:
        t.check_grad()
