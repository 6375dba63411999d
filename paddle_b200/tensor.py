"""paddle_b200.Tensor — the eager tensor.

Parity target: ``paddle.Tensor`` (reference: python/paddle/tensor/tensor.prototype.pyi,
paddle/fluid/pybind/eager_method.cc, eager_properties.cc).

Design: a ``torch.Tensor`` subclass.  Storage, autograd graph, CUDA caching
allocator and stream semantics are PyTorch's; the paddle surface
(``stop_gradient``, list-valued ``shape``, ``axis=`` keyword style, paddle
method names) lives here.  Every result of a torch op on a ``Tensor`` is
re-wrapped as ``Tensor`` (never as ``Parameter``).
"""
from __future__ import annotations

import numpy as np
import torch

from .framework import dtype as _dt
from .framework import place as _place
from .framework import recording as _rec

_NOWRAP = None


def _nowrap():
    global _NOWRAP
    if _NOWRAP is None:
        from torch.overrides import get_default_nowrap_functions

        _NOWRAP = set(get_default_nowrap_functions())
    return _NOWRAP


def _convert(ret):
    if isinstance(ret, torch.Tensor):
        if not isinstance(ret, Tensor):
            return ret.as_subclass(Tensor)
        return ret
    if isinstance(ret, (tuple, list)):
        # keep namedtuple-ness (torch.return_types) out: plain containers are enough for paddle API
        return type(ret)(_convert(r) for r in ret) if type(ret) in (tuple, list) else tuple(_convert(r) for r in ret)
    return ret


class _BoolCall(int):
    """0 / 1 that can also be called to get the bool (attribute-style and method-style spellings of the same predicate)."""

    def __call__(self):
        return bool(self)

    def __repr__(self):
        return "True" if self else "False"

    __str__ = __repr__


class Tensor(torch.Tensor):
    """Eager tensor with paddle semantics."""

    # ---- torch integration -------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        with torch._C.DisableTorchFunctionSubclass():
            ret = func(*args, **(kwargs or {}))
        if func in _nowrap():
            return ret
        return _convert(ret)

    def __reduce_ex__(self, proto):
        # generic pickling (multiprocessing, jit.save of a Layer) keeps the tensor type; paddle.save uses its own
        # dispatch table to write the reference's (name, ndarray) form (framework/io.py:_pickle_dump)
        return (_rebuild_tensor, (type(self), self.detach().cpu().as_subclass(torch.Tensor), self.requires_grad, dict(self.__dict__)))

    def __deepcopy__(self, memo):
        with torch.no_grad():
            new = self.as_subclass(torch.Tensor).clone().as_subclass(type(self))
        new.requires_grad_(self.requires_grad)
        for k, v in self.__dict__.items():
            new.__dict__[k] = v
        memo[id(self)] = new
        return new

    # ---- paddle attributes -------------------------------------------------
    @property
    def stop_gradient(self) -> bool:
        return not self.requires_grad

    @stop_gradient.setter
    def stop_gradient(self, v: bool):
        v = bool(v)
        if v:
            if self.requires_grad:
                if self.is_leaf:
                    self.requires_grad_(False)
                else:
                    self.detach_()
        else:
            if not self.requires_grad and (self.is_floating_point() or self.is_complex()):
                self.requires_grad_(True)

    @property
    def shape(self):  # paddle returns list
        return list(self.size())

    @property
    def place(self):
        return _place.place_of(self)

    @property
    def name(self):
        n = self.__dict__.get("_pd_name")
        if n is None:
            from .framework import unique_name

            n = unique_name.generate("generated_tensor")
            self.__dict__["_pd_name"] = n
        return n

    @name.setter
    def name(self, v):
        self.__dict__["_pd_name"] = v

    @property
    def persistable(self):
        return self.__dict__.get("_pd_persistable", False)

    @persistable.setter
    def persistable(self, v):
        self.__dict__["_pd_persistable"] = bool(v)

    @property
    def size(self):  # paddle: number of elements (int). torch's .size() stays callable via _SizeProxy
        return _SizeProxy(self)

    @property
    def T(self):
        nd = self.dim()
        return self.permute(*reversed(range(nd))) if nd != 2 else torch.Tensor.t(self)

    @property
    def mT(self):
        return torch.transpose(self, -1, -2)

    @property
    def grad(self):
        g = torch.Tensor.grad.__get__(self)
        if g is not None and not isinstance(g, Tensor):
            g = g.as_subclass(Tensor)
        return g

    @grad.setter
    def grad(self, v):
        torch.Tensor.grad.__set__(self, None if v is None else v.as_subclass(torch.Tensor))

    @property
    def is_leaf(self):
        return torch.Tensor.is_leaf.__get__(self)

    @property
    def is_sparse(self):
        """paddle spells it `x.is_sparse()`, torch `x.is_sparse`: the value answers both (truthy int that is also callable)."""
        return _BoolCall(torch.Tensor.layout.__get__(self) != torch.strided)

    def is_dense(self):
        return self.layout == torch.strided

    def is_dist(self):
        return self.__dict__.get("_pd_dist_attr") is not None

    def is_sparse_coo(self):
        return self.layout == torch.sparse_coo

    def is_sparse_csr(self):
        return self.layout == torch.sparse_csr

    def _is_initialized(self):
        return True

    # ---- conversion --------------------------------------------------------
    def _no_static_value(self, what):
        """A value of a program that is being built has no data yet: a host read would hand the PLACEHOLDER's zeros to python code (an `if`, a loop
        bound, a shape) and freeze them into the program.  The reference raises here too."""
        prog = _rec.current[0]
        if prog is not None and _rec._inside[0] == 0 and id(self) in prog._vids:
            raise RuntimeError(f"{what} of a program variable while the program is being built: it has no value yet (fetch it through Executor.run, "
                               "or use static.nn.cond / while_loop for value-dependent control flow)")

    def numpy(self):
        if _rec.current[0] is not None:
            self._no_static_value(".numpy()")
        t = self.detach().as_subclass(torch.Tensor)
        if t.device.type != "cpu":
            t = t.cpu()
        if t.is_conj() or t.is_neg():   # lazy conj / neg views (fft, linalg) have no numpy form
            t = t.resolve_conj().resolve_neg()
        if t.dtype == torch.bfloat16:
            return t.view(torch.int16).numpy().view(np.uint16)
        if t.dtype in (torch.float8_e4m3fn, torch.float8_e5m2):
            return t.view(torch.uint8).numpy()
        return t.numpy()

    # ---- indexing: negative slice steps (numpy / paddle semantics; torch itself rejects them) ----------------------------------
    @staticmethod
    def _has_neg_step(idx):
        if type(idx) is slice:
            return idx.step is not None and idx.step < 0
        if type(idx) is tuple:
            for i in idx:
                if type(i) is slice and i.step is not None and i.step < 0:
                    return True
        return False

    def _dims_of_index(self, idx):
        """[(position in idx, tensor dim)] of the slice entries of a basic / mixed index."""
        items = list(idx)
        n_consumed = 0
        for i in items:
            if i is None or i is Ellipsis:
                continue
            n_consumed += i.dim() if (isinstance(i, torch.Tensor) and i.dtype == torch.bool) else 1
        out, d = [], 0
        for pos, i in enumerate(items):
            if i is None:
                continue
            if i is Ellipsis:
                d += self.dim() - n_consumed
                continue
            if type(i) is slice:
                out.append((pos, d))
            d += i.dim() if (isinstance(i, torch.Tensor) and i.dtype == torch.bool) else 1
        return out

    def __getitem__(self, idx):
        if not Tensor._has_neg_step(idx):
            return torch.Tensor.__getitem__(self, idx)
        items = list(idx) if type(idx) is tuple else [idx]
        t, flips = self, []
        for pos, d in self._dims_of_index(items):
            sl = items[pos]
            if sl.step is not None and sl.step < 0:
                size = int(torch.Tensor.size(self, d))
                start, stop, step = sl.indices(size)
                n = len(range(start, stop, step))
                flips.append(d)
                # the same elements in a flipped axis: i' = size - 1 - i, visited with a positive step
                items[pos] = slice(size - 1 - start, size - 1 - start + n * (-step), -step) if n else slice(0, 0, 1)
        if flips:
            t = torch.flip(self, flips)
        return torch.Tensor.__getitem__(t, tuple(items))

    def __setitem__(self, idx, value):
        if not Tensor._has_neg_step(idx):
            return torch.Tensor.__setitem__(self, idx, value)
        items = list(idx) if type(idx) is tuple else [idx]
        if any(isinstance(i, (torch.Tensor, list)) for i in items):
            raise NotImplementedError("assignment through a negative-step slice combined with tensor / list indices")
        neg = [(pos, d) for pos, d in self._dims_of_index(items) if items[pos].step is not None and items[pos].step < 0]
        if len(neg) > 1:
            raise NotImplementedError("assignment through more than one negative-step slice")
        pos, d = neg[0]
        size = int(torch.Tensor.size(self, d))
        items[pos] = torch.arange(*items[pos].indices(size), device=self.device)      # one index vector keeps basic-index result layout
        return torch.Tensor.__setitem__(self, tuple(items), value)

    # ---- LoD (level-of-detail offsets of packed variable-length sequences; used by static.nn.sequence_* ops) ----
    def lod(self):
        return [list(o) for o in self.__dict__.get("_lod", [])]

    def set_lod(self, lod):
        self.__dict__["_lod"] = [[int(v) for v in o] for o in lod]
        return self

    def recursive_sequence_lengths(self):
        return [[b - a for a, b in zip(o[:-1], o[1:])] for o in self.__dict__.get("_lod", [])]

    def set_recursive_sequence_lengths(self, lens):
        lod = []
        for l in lens:
            o = [0]
            for n in l:
                o.append(o[-1] + int(n))
            lod.append(o)
        self.__dict__["_lod"] = lod
        return self

    def has_valid_recursive_sequence_lengths(self):
        lod = self.__dict__.get("_lod", [])
        return bool(lod) and lod[-1][-1] == self.shape[0] and all(a <= b for o in lod for a, b in zip(o[:-1], o[1:]))

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        if a.dtype == np.uint16 and self.dtype == torch.bfloat16:
            a = self.detach().float().cpu().as_subclass(torch.Tensor).numpy()
        return a.astype(dtype) if dtype is not None else a

    def astype(self, dtype):
        return self.to(_dt.convert_dtype(dtype))

    cast = astype

    def item(self, *idx):
        if _rec.current[0] is not None:
            self._no_static_value(".item()")
        if idx:
            return torch.Tensor.item(self.reshape(-1)[idx[0]] if len(idx) == 1 else self[idx])
        return torch.Tensor.item(self)

    def cpu(self):
        return self.to("cpu")

    def cuda(self, device_id=None, blocking=True):
        dev = torch.device("cuda", device_id if device_id is not None else torch.cuda.current_device())
        return self.to(dev, non_blocking=not blocking)

    def pin_memory(self, *a, **k):
        return torch.Tensor.pin_memory(self)

    def to(self, *args, **kwargs):
        # paddle: to(device) / to(dtype) / to(device, dtype, blocking) / to(other)
        blocking = kwargs.pop("blocking", None)
        new_args = []
        for a in args:
            if isinstance(a, (_place.Place,)):
                new_args.append(a.to_torch())
            elif isinstance(a, str):
                try:
                    new_args.append(_dt.convert_dtype(a))
                except TypeError:
                    new_args.append(_place.to_torch_device(a))
            elif isinstance(a, bool) and blocking is None and len(new_args) >= 1:
                blocking = a
            else:
                new_args.append(a)
        if "device" in kwargs:
            kwargs["device"] = _place.to_torch_device(kwargs["device"])
        if "dtype" in kwargs:
            kwargs["dtype"] = _dt.convert_dtype(kwargs["dtype"])
        if blocking is not None:
            kwargs["non_blocking"] = not blocking
        return torch.Tensor.to(self, *new_args, **kwargs)

    def _copy_to(self, place, blocking=True):
        return self.to(_place.to_torch_device(place), non_blocking=not blocking)

    def _to(self, device=None, dtype=None, blocking=None):
        out = self
        if device is not None:
            out = out.to(_place.to_torch_device(device))
        if dtype is not None:
            out = out.astype(dtype)
        return out

    def value(self):
        return self

    def get_tensor(self):
        return self

    # ---- autograd ----------------------------------------------------------
    def backward(self, grad_tensor=None, retain_graph=False):
        torch.autograd.backward(self, grad_tensors=grad_tensor, retain_graph=retain_graph)

    def clear_grad(self, set_to_zero=False):
        if set_to_zero and torch.Tensor.grad.__get__(self) is not None:
            torch.Tensor.grad.__get__(self).zero_()
        else:
            torch.Tensor.grad.__set__(self, None)

    clear_gradient = clear_grad

    def gradient(self):
        g = self.grad
        return None if g is None else g.numpy()

    def detach(self):
        out = torch.Tensor.detach(self)
        return out

    def register_hook(self, hook):
        return torch.Tensor.register_hook(self, hook)

    def set_value(self, value):
        """Immediate assignment (not an op of a program being built): the value has to be concrete."""
        if _rec.current[0] is not None and isinstance(value, torch.Tensor) and _rec._inside[0] == 0 and id(value) in _rec.current[0]._vids:
            raise RuntimeError("set_value needs a concrete value, got a value of the program under construction (it has no data yet): pass a numpy array / a tensor "
                               "built outside program_guard, or assign through the scope after the program ran")
        with torch.no_grad(), torch._C.DisableTorchFunction():          # an action on this tensor now, never a recorded op
            if not isinstance(value, torch.Tensor):
                value = torch.as_tensor(np.asarray(value))
            torch.Tensor.copy_(self, value.to(device=self.device).reshape(self.size()))
        return self

    def copy_(self, other, blocking=True):
        return torch.Tensor.copy_(self, other, non_blocking=not blocking)

    def get_strides(self):
        return list(self.stride())

    def element_size(self):
        return torch.Tensor.element_size(self)

    # ---- python protocol ---------------------------------------------------
    def __repr__(self):
        try:
            body = np.array2string(self.detach().float().cpu().as_subclass(torch.Tensor).numpy()
                                   if self.dtype in (torch.bfloat16, torch.float16) else
                                   self.detach().cpu().as_subclass(torch.Tensor).numpy(),
                                   separator=", ", prefix="       ")
        except Exception:
            body = "<unprintable>"
        return (f"Tensor(shape={self.shape}, dtype={_dt.dtype_name(self.dtype)}, place={self.place}, "
                f"stop_gradient={self.stop_gradient},\n       {body})")

    __str__ = __repr__

    def __format__(self, spec):
        if self.dim() == 0:
            return format(self.item(), spec)
        return repr(self)

    def __hash__(self):
        return id(self)

    def __len__(self):
        if self.dim() == 0:
            raise TypeError("len() of a 0-D tensor")
        return self.size(0)

    def __bool__(self):
        if _rec.current[0] is not None:
            self._no_static_value("bool()")
        return bool(torch.Tensor.item(self)) if self.numel() == 1 else torch.Tensor.__bool__(self)

    def __float__(self):
        if _rec.current[0] is not None:
            self._no_static_value("float()")
        return torch.Tensor.__float__(self)

    def __int__(self):
        if _rec.current[0] is not None:
            self._no_static_value("int()")
        return torch.Tensor.__int__(self)

    def __index__(self):
        if _rec.current[0] is not None:
            self._no_static_value("an integer index")
        return torch.Tensor.__index__(self)

    def tolist(self):
        if _rec.current[0] is not None:
            self._no_static_value(".tolist()")
        return torch.Tensor.tolist(self)

    def dim(self):
        return torch.Tensor.dim(self)

    ndimension = dim

    def rank(self):
        return self.dim()


class _SizeProxy(int):
    """``x.size`` is an int in paddle; keep ``x.size()`` / ``x.size(0)`` (torch style) callable."""

    def __new__(cls, t):
        obj = int.__new__(cls, torch.Tensor.numel(t))
        obj._t = t
        return obj

    def __call__(self, *a):
        return torch.Tensor.size(self._t, *a)


class Parameter(Tensor):
    """Trainable parameter. Parity: python/paddle/base/framework.py:EagerParamBase."""

    def __new__(cls, data, trainable=True, name=None):
        if isinstance(data, torch.Tensor):
            base = data.detach().as_subclass(torch.Tensor)
        else:
            base = torch.as_tensor(data)
        p = torch.Tensor._make_subclass(cls, base, trainable and (base.is_floating_point() or base.is_complex()))
        return p

    def __init__(self, data, trainable=True, name=None):
        from .framework import unique_name

        self.__dict__["_pd_name"] = name or unique_name.generate("param")
        self.__dict__["_pd_persistable"] = True
        self.__dict__["_pd_trainable"] = trainable
        self.optimize_attr = {"learning_rate": 1.0}
        self.regularizer = None
        self.do_model_average = None
        self.need_clip = True
        self.is_distributed = False

    @property
    def trainable(self):
        return not self.stop_gradient

    @trainable.setter
    def trainable(self, v):
        self.stop_gradient = not v

    def __deepcopy__(self, memo):
        new = Parameter(self.detach().clone(), trainable=self.requires_grad, name=None)
        for k, v in self.__dict__.items():
            if k != "_pd_name":
                new.__dict__[k] = v
        memo[id(self)] = new
        return new

EagerParamBase = Parameter


def _rebuild_tensor(cls, base, requires_grad, attrs):
    if issubclass(cls, Parameter):
        t = Parameter(base, trainable=requires_grad, name=attrs.get("_pd_name"))
    else:
        t = base.as_subclass(cls)
        if requires_grad:
            t.requires_grad_(True)
    for k, v in attrs.items():
        if k not in ("_arena_grad",):
            t.__dict__[k] = v
    return t


def _np_to_torch(a: np.ndarray):
    if a.dtype == np.uint16:  # paddle stores bf16 as uint16
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
    if not a.flags.writeable or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a).copy()
    return torch.from_numpy(a)


def to_tensor(data, dtype=None, place=None, stop_gradient=True):
    """paddle.to_tensor. Parity: python/paddle/tensor/creation.py:to_tensor."""
    dtype = _dt.convert_dtype(dtype)
    dev = _place.to_torch_device(place)
    if isinstance(data, torch.Tensor):
        t = data.detach().as_subclass(torch.Tensor)
        t = t.to(device=dev, dtype=dtype if dtype is not None else t.dtype, copy=True)
    else:
        if isinstance(data, (list, tuple)) and any(isinstance(d, torch.Tensor) for d in _flatten(data)):
            t = torch.stack([to_tensor(d).as_subclass(torch.Tensor) for d in data])
        elif isinstance(data, np.ndarray):
            t = _np_to_torch(data)
        else:
            a = np.asarray(data)
            if a.dtype == np.float64 and dtype is None and not isinstance(data, np.generic):
                a = a.astype(_dt.to_numpy_dtype(_dt.default_dtype())) if _dt.default_dtype() != torch.bfloat16 else a.astype(np.float32)
            t = _np_to_torch(a)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        pinned = isinstance(place, _place.CUDAPinnedPlace)
        if t.device != dev:
            t = t.to(dev)
        elif isinstance(data, np.ndarray) and data.flags.writeable and data.flags.c_contiguous and data.dtype != np.uint16:
            t = t.clone()  # paddle copies
        if pinned and torch.cuda.is_available():
            t = t.pin_memory()
    out = t.as_subclass(Tensor)
    if not stop_gradient and (out.is_floating_point() or out.is_complex()):
        out.requires_grad_(True)
    return out


def _flatten(x):
    for i in x:
        if isinstance(i, (list, tuple)):
            yield from _flatten(i)
        else:
            yield i


def as_tensor(x, dtype=None, device=None) -> Tensor:
    """Internal: cheap coercion without copy when already a tensor."""
    if isinstance(x, Tensor):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    if isinstance(x, torch.Tensor):
        return x.as_subclass(Tensor)
    return to_tensor(x, dtype=dtype, place=device)


def is_tensor(x) -> bool:
    return isinstance(x, torch.Tensor)
