"""nn.Layer base class. Parity: python/paddle/nn/layer/layers.py:Layer, python/paddle/base/param_attr.py."""
from __future__ import annotations

import collections
import re
from collections import OrderedDict

import numpy as np
import torch

from ..framework import dtype as _dt
from ..framework import place as _place
from ..framework import unique_name
from ..tensor import Parameter, Tensor


class ParamAttr:
    """Parity: python/paddle/base/param_attr.py:ParamAttr."""

    def __init__(self, name=None, initializer=None, learning_rate=1.0, regularizer=None, trainable=True,
                 do_model_average=True, need_clip=True):
        self.name, self.initializer, self.learning_rate = name, initializer, learning_rate
        self.regularizer, self.trainable = regularizer, trainable
        self.do_model_average, self.need_clip = do_model_average, need_clip

    @staticmethod
    def _to_attr(arg):
        if arg is None:
            return ParamAttr()
        if isinstance(arg, ParamAttr):
            return arg
        if isinstance(arg, str):
            return ParamAttr(name=arg)
        if isinstance(arg, bool):
            return ParamAttr() if arg else False
        from . import initializer as I

        if isinstance(arg, I.Initializer):
            return ParamAttr(initializer=arg)
        raise TypeError(f"cannot convert {type(arg)} to ParamAttr")


_lazy_init = [False]


def _make_parameter(shape, dtype, attr=None, is_bias=False, default_initializer=None, name=None, prefix="param"):
    from . import initializer as I

    attr = ParamAttr._to_attr(attr)
    if attr is False:
        return None
    dtype = _dt.convert_dtype(dtype) or _dt.default_dtype()
    dev = _place.to_torch_device(None)
    data = torch.empty([int(s) for s in shape], dtype=dtype, device=dev)
    p = Parameter(data, trainable=attr.trainable, name=attr.name or name or unique_name.generate(prefix))
    init = attr.initializer or I._global_initializer(is_bias) or default_initializer
    if init is None:
        init = I.Constant(0.0) if is_bias else I.XavierUniform()
    with torch.no_grad():
        init(p)
    p.optimize_attr = {"learning_rate": attr.learning_rate}
    p.regularizer = attr.regularizer
    p.do_model_average = attr.do_model_average
    p.need_clip = attr.need_clip
    return p


class HookRemoveHelper:
    _next = [0]

    def __init__(self, hooks):
        self._hooks = hooks
        self._id = HookRemoveHelper._next[0]
        HookRemoveHelper._next[0] += 1

    def remove(self):
        self._hooks.pop(self._id, None)


def _camel_to_snake(n):
    return re.sub(r"(?<!^)(?=[A-Z])", "_", n).lower()


class Layer:
    """Base class of all network layers (dygraph)."""

    def __init__(self, name_scope=None, dtype="float32"):
        d = self.__dict__
        d["training"] = True
        d["_parameters"] = OrderedDict()
        d["_buffers"] = OrderedDict()
        d["_non_persistable_buffer_names_set"] = set()
        d["_sub_layers"] = OrderedDict()
        d["_forward_pre_hooks"] = OrderedDict()
        d["_forward_post_hooks"] = OrderedDict()
        # like the reference's built-in layers (LayerHelper.get_default_dtype) parameters follow paddle.set_default_dtype
        d["_dtype"] = _dt.default_dtype() if dtype in (None, "float32") else _dt.convert_dtype(dtype)
        base = name_scope or _camel_to_snake(self.__class__.__name__)
        d["_full_name"] = unique_name.generate(base)
        d["_helper_w"] = 0
        d["_helper_b"] = 0
        d["_casted_by_pure_fp16"] = False

    # ---- construction helpers ----------------------------------------------
    def full_name(self):
        return self._full_name

    def create_parameter(self, shape, attr=None, dtype=None, is_bias=False, default_initializer=None):
        if attr is False:
            return None
        suffix = "b" if is_bias else "w"
        cnt_key = "_helper_b" if is_bias else "_helper_w"
        idx = self.__dict__[cnt_key]
        self.__dict__[cnt_key] = idx + 1
        name = f"{self._full_name}.{suffix}_{idx}"
        return _make_parameter(shape, dtype or self._dtype, attr=attr, is_bias=is_bias,
                               default_initializer=default_initializer, name=name)

    def create_variable(self, name=None, persistable=None, dtype=None):
        t = torch.empty(0, dtype=_dt.convert_dtype(dtype) or self._dtype).as_subclass(Tensor)
        if name:
            t.name = name
        t.persistable = bool(persistable)
        return t

    create_tensor = create_variable

    def add_parameter(self, name, parameter):
        if parameter is not None and not isinstance(parameter, Parameter):
            raise TypeError("add_parameter expects a Parameter or None")
        self._parameters[name] = parameter
        return parameter

    def add_sublayer(self, name, sublayer):
        if sublayer is not None and not isinstance(sublayer, Layer):
            raise TypeError("add_sublayer expects a Layer or None")
        self._sub_layers[str(name)] = sublayer
        return sublayer

    def register_buffer(self, name, tensor, persistable=True):
        if tensor is not None and not isinstance(tensor, torch.Tensor):
            raise TypeError("register_buffer expects a Tensor or None")
        if tensor is not None and tensor.device.type == "cpu":
            from ..framework import place as _place

            d = _place.to_torch_device(None)
            if d.type != "cpu":            # buffers follow the current device like parameters do (paddle.set_device)
                tensor = tensor.to(d)
        if tensor is not None and not isinstance(tensor, Tensor):
            tensor = tensor.as_subclass(Tensor)
        self._buffers[name] = tensor
        if persistable:
            self._non_persistable_buffer_names_set.discard(name)
        else:
            self._non_persistable_buffer_names_set.add(name)

    # ---- attribute protocol ------------------------------------------------
    def __setattr__(self, name, value):
        d = self.__dict__
        if isinstance(value, Parameter):
            if "_parameters" not in d:
                raise RuntimeError("call super().__init__() before assigning parameters")
            for store in (d["_sub_layers"], d["_buffers"]):
                store.pop(name, None)
            d.pop(name, None)
            d["_parameters"][name] = value
        elif isinstance(value, Layer):
            if "_sub_layers" not in d:
                raise RuntimeError("call super().__init__() before assigning sublayers")
            for store in (d["_parameters"], d["_buffers"]):
                store.pop(name, None)
            d.pop(name, None)
            d["_sub_layers"][name] = value
        elif "_parameters" in d and name in d["_parameters"]:
            if value is not None:
                raise TypeError(f"cannot assign non-Parameter to parameter '{name}'")
            d["_parameters"][name] = None
        elif "_sub_layers" in d and name in d["_sub_layers"]:
            if value is None:
                d["_sub_layers"][name] = None
            else:
                del d["_sub_layers"][name]
                object.__setattr__(self, name, value)
        elif "_buffers" in d and name in d["_buffers"]:
            if value is not None and not isinstance(value, torch.Tensor):
                raise TypeError(f"cannot assign non-Tensor to buffer '{name}'")
            d["_buffers"][name] = value if value is None or isinstance(value, Tensor) else value.as_subclass(Tensor)
        else:
            object.__setattr__(self, name, value)

    def __getattr__(self, name):
        d = self.__dict__
        for store in ("_parameters", "_sub_layers", "_buffers"):
            s = d.get(store)
            if s is not None and name in s:
                return s[name]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def __delattr__(self, name):
        for store in ("_parameters", "_sub_layers", "_buffers"):
            if name in self.__dict__.get(store, {}):
                del self.__dict__[store][name]
                return
        object.__delattr__(self, name)

    def __dir__(self):
        return list(super().__dir__()) + list(self._parameters) + list(self._sub_layers) + list(self._buffers)

    # ---- call --------------------------------------------------------------
    def forward(self, *inputs, **kwargs):
        raise NotImplementedError

    def __call__(self, *inputs, **kwargs):
        if self._forward_pre_hooks:
            for hook in list(self._forward_pre_hooks.values()):
                r = hook(self, inputs)
                if r is not None:
                    inputs = r if isinstance(r, tuple) else (r,)
        out = self.forward(*inputs, **kwargs)
        if self._forward_post_hooks:
            for hook in list(self._forward_post_hooks.values()):
                r = hook(self, inputs, out)
                if r is not None:
                    out = r
        return out

    def register_forward_pre_hook(self, hook):
        h = HookRemoveHelper(self._forward_pre_hooks)
        self._forward_pre_hooks[h._id] = hook
        return h

    def register_forward_post_hook(self, hook):
        h = HookRemoveHelper(self._forward_post_hooks)
        self._forward_post_hooks[h._id] = hook
        return h

    # ---- traversal ---------------------------------------------------------
    def named_sublayers(self, prefix="", include_self=False, layers_set=None, remove_duplicate=True):
        if layers_set is None:
            layers_set = set()
        if include_self and (not remove_duplicate or id(self) not in layers_set):
            layers_set.add(id(self))
            yield prefix, self
        for k, l in self._sub_layers.items():
            if l is None or (remove_duplicate and id(l) in layers_set):
                continue
            p = prefix + ("." if prefix else "") + k
            layers_set.add(id(l))
            yield p, l
            yield from l.named_sublayers(prefix=p, include_self=False, layers_set=layers_set, remove_duplicate=remove_duplicate)

    def sublayers(self, include_self=False):
        return [l for _, l in self.named_sublayers(include_self=include_self)]

    def children(self):
        return (l for l in self._sub_layers.values() if l is not None)

    def named_children(self):
        return ((k, l) for k, l in self._sub_layers.items() if l is not None)

    def named_parameters(self, prefix="", include_sublayers=True, remove_duplicate=True):
        seen = set()
        layers = self.named_sublayers(prefix=prefix, include_self=True, remove_duplicate=remove_duplicate) if include_sublayers else [(prefix, self)]
        for lp, l in layers:
            for k, p in l._parameters.items():
                if p is None or (remove_duplicate and id(p) in seen):
                    continue
                seen.add(id(p))
                yield lp + ("." if lp else "") + k, p

    def parameters(self, include_sublayers=True):
        return [p for _, p in self.named_parameters(include_sublayers=include_sublayers)]

    def named_buffers(self, prefix="", include_sublayers=True, remove_duplicate=True):
        seen = set()
        layers = self.named_sublayers(prefix=prefix, include_self=True, remove_duplicate=remove_duplicate) if include_sublayers else [(prefix, self)]
        for lp, l in layers:
            for k, b in l._buffers.items():
                if b is None or (remove_duplicate and id(b) in seen):
                    continue
                seen.add(id(b))
                yield lp + ("." if lp else "") + k, b

    def buffers(self, include_sublayers=True):
        return [b for _, b in self.named_buffers(include_sublayers=include_sublayers)]

    def apply(self, fn):
        for l in self.children():
            l.apply(fn)
        fn(self)
        return self

    # ---- mode / dtype / device ---------------------------------------------
    def train(self):
        for l in self.sublayers(include_self=True):
            l.__dict__["training"] = True
        return self

    def eval(self):
        for l in self.sublayers(include_self=True):
            l.__dict__["training"] = False
        return self

    def _apply_tensors(self, fn):
        for l in self.sublayers(include_self=True):
            for k, p in l._parameters.items():
                if p is not None:
                    new = fn(p)
                    if new is not p:
                        with torch.no_grad():
                            p.data = new.as_subclass(torch.Tensor)
                        g = torch.Tensor.grad.__get__(p)
                        if g is not None:
                            torch.Tensor.grad.__set__(p, fn(g.as_subclass(Tensor)).as_subclass(torch.Tensor))
            for k, b in l._buffers.items():
                if b is not None:
                    l._buffers[k] = fn(b)
        return self

    def to(self, device=None, dtype=None, blocking=None):
        dev = _place.to_torch_device(device) if device is not None else None
        dty = _dt.convert_dtype(dtype)

        def fn(t):
            d = dty if (dty is not None and t.is_floating_point()) else None
            if dev is None and d is None:
                return t
            return torch.Tensor.to(t, device=dev if dev is not None else t.device, dtype=d if d is not None else t.dtype)

        if dty is not None:
            for l in self.sublayers(include_self=True):
                l.__dict__["_dtype"] = dty
        return self._apply_tensors(fn)

    def astype(self, dtype):
        return self.to(dtype=dtype)

    def float(self, excluded_layers=None):
        return self._cast_floating(torch.float32, excluded_layers)

    def float16(self, excluded_layers=None):
        return self._cast_floating(torch.float16, excluded_layers)

    def bfloat16(self, excluded_layers=None):
        return self._cast_floating(torch.bfloat16, excluded_layers)

    def _cast_floating(self, dtype, excluded_layers=None):
        excluded = tuple(excluded_layers) if isinstance(excluded_layers, (list, tuple)) else ((excluded_layers,) if excluded_layers else ())
        for l in self.sublayers(include_self=True):
            if excluded and isinstance(l, excluded):
                continue
            for k, p in l._parameters.items():
                if p is not None and p.is_floating_point():
                    with torch.no_grad():
                        p.data = p.as_subclass(torch.Tensor).to(dtype)
            for k, b in l._buffers.items():
                if b is not None and b.is_floating_point():
                    l._buffers[k] = torch.Tensor.to(b, dtype)
            l.__dict__["_dtype"] = dtype
        return self

    def cuda(self, device_id=None):
        return self.to(torch.device("cuda", device_id if device_id is not None else torch.cuda.current_device()))

    def cpu(self):
        return self.to("cpu")

    def clear_gradients(self, set_to_zero=False):
        for p in self.parameters():
            p.clear_grad(set_to_zero)

    # ---- state dict --------------------------------------------------------
    def state_dict(self, destination=None, include_sublayers=True, structured_name_prefix="", use_hook=True,
                   keep_vars=True):
        dest = destination if destination is not None else OrderedDict()
        for name, p in self.named_parameters(prefix=structured_name_prefix.rstrip("."), include_sublayers=include_sublayers):
            dest[name] = p
        layers = self.named_sublayers(prefix=structured_name_prefix.rstrip("."), include_self=True) if include_sublayers else [("", self)]
        seen = set()
        for lp, l in layers:
            for k, b in l._buffers.items():
                if b is None or k in l._non_persistable_buffer_names_set or id(b) in seen:
                    continue
                seen.add(id(b))
                dest[lp + ("." if lp else "") + k] = b
        if use_hook:
            for hook in self.__dict__.get("_state_dict_hooks", {}).values():
                res = hook(dest)
                if res is not None:
                    dest = res
        return dest

    def register_state_dict_hook(self, hook):
        """hook(state_dict) -> state_dict | None, applied to the result of state_dict(). Parity: layers.py:register_state_dict_hook."""
        hooks = self.__dict__.setdefault("_state_dict_hooks", OrderedDict())
        h = HookRemoveHelper(hooks)
        hooks[h._id] = hook
        return h

    def backward(self, *inputs):
        raise ValueError("Layer shouldn't implement backward")

    def to_static_state_dict(self, *a, **k):
        return self.state_dict(*a, **k)

    def set_state_dict(self, state_dict, use_structured_name=True):
        """Returns (missing_keys, unexpected_keys). Parity: layers.py:set_state_dict."""
        own = self.state_dict()
        if not use_structured_name:
            by_name = {v.name: k for k, v in own.items()}
            state_dict = {by_name.get(k, k): v for k, v in state_dict.items()}
        missing, unexpected = [], []
        for k, t in own.items():
            if k not in state_dict:
                missing.append(k)
                continue
            v = state_dict[k]
            if isinstance(v, tuple) and len(v) == 2 and isinstance(v[1], np.ndarray):
                v = v[1]
            if isinstance(v, np.ndarray):
                from ..tensor import _np_to_torch

                v = _np_to_torch(v)
            elif not isinstance(v, torch.Tensor):
                v = torch.as_tensor(np.asarray(v))
            if list(v.shape) != list(t.size()):
                raise ValueError(f"shape mismatch for '{k}': checkpoint {list(v.shape)} vs layer {list(t.size())}")
            with torch.no_grad():
                torch.Tensor.copy_(t, v.to(device=t.device, dtype=t.dtype))
        for k in state_dict:
            if k not in own and k != "StructuredToParameterName@@":
                unexpected.append(k)
        return missing, unexpected

    set_dict = set_state_dict
    load_dict = set_state_dict

    # ---- misc --------------------------------------------------------------
    def extra_repr(self):
        return ""

    def __repr__(self):
        lines = []
        extra = self.extra_repr()
        for k, l in self._sub_layers.items():
            s = repr(l).split("\n")
            lines.append(f"({k}): " + s[0])
            lines.extend(s[1:])
        main = self.__class__.__name__ + "("
        if extra and not lines:
            return main + extra + ")"
        if extra:
            lines.insert(0, extra)
        if lines:
            main += "\n  " + "\n  ".join(lines) + "\n"
        return main + ")"


# torch interop used by a few subsystems (e.g. functional_call-style utilities)
collections.abc.Callable.register(Layer)
