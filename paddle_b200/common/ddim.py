"""DDim: the dimension vector used by infer-meta / kernels. Parity: paddle/common/ddim.{h,cc} (make_ddim, product, slice_ddim,
flatten_to_2d, stride, vectorize); -1 marks a dimension unknown until run time."""
from __future__ import annotations

from typing import Iterable, Sequence


class DDim:
    __slots__ = ("_d",)
    MAX_RANK = 9

    def __init__(self, dims: Iterable[int] = ()):
        d = tuple(int(x) for x in dims)
        if len(d) > self.MAX_RANK:
            from .errors import InvalidArgumentError

            raise InvalidArgumentError(f"DDim supports rank <= {self.MAX_RANK}, got {len(d)}")
        self._d = d

    def size(self):
        return len(self._d)

    def __len__(self):
        return len(self._d)

    def __iter__(self):
        return iter(self._d)

    def __getitem__(self, i):
        r = self._d[i]
        return DDim(r) if isinstance(i, slice) else r

    def at(self, i):
        if not -len(self._d) <= i < len(self._d):
            from .errors import OutOfRangeError

            raise OutOfRangeError(f"DDim index {i} out of range for rank {len(self._d)}")
        return self._d[i]

    def __eq__(self, o):
        return tuple(o) == self._d if isinstance(o, (DDim, tuple, list)) else NotImplemented

    def __hash__(self):
        return hash(self._d)

    def __repr__(self):
        return "DDim([" + ", ".join(map(str, self._d)) + "])"

    def to_list(self):
        return list(self._d)

    def is_dynamic(self):
        return any(x < 0 for x in self._d)

    def reshape(self, shape: Sequence[int]):
        """Resolve one -1 / 0 (copy) entry like the reshape infer-meta."""
        shape = list(shape)
        known = 1
        unk = None
        for i, s in enumerate(shape):
            if s == 0:
                shape[i] = self._d[i]
            if shape[i] == -1:
                unk = i
            else:
                known *= shape[i]
        if unk is not None:
            total = product(self)
            shape[unk] = total // known if total >= 0 and known > 0 else -1
        return DDim(shape)


def make_ddim(dims):
    return dims if isinstance(dims, DDim) else DDim(dims)


def vectorize(d):
    return list(make_ddim(d))


def product(d):
    p = 1
    for x in make_ddim(d):
        if x < 0:
            return -1
        p *= x
    return p


def slice_ddim(d, begin, end):
    return make_ddim(d)[begin:end]


def flatten_to_2d(d, num_col_dims):
    d = make_ddim(d)
    return DDim([product(d[:num_col_dims]), product(d[num_col_dims:])])


def flatten_to_1d(d):
    return DDim([product(d)])


def stride(d):
    """Row-major (contiguous) strides in elements."""
    d = vectorize(d)
    out, acc = [0] * len(d), 1
    for i in range(len(d) - 1, -1, -1):
        out[i] = acc
        acc *= max(d[i], 1)
    return DDim(out)


def stride_numel(d):
    """numel of every suffix: stride_numel[i] = prod(d[i:])."""
    d = vectorize(d)
    out, acc = [0] * len(d), 1
    for i in range(len(d) - 1, -1, -1):
        acc *= d[i]
        out[i] = acc
    return DDim(out)
