"""Optimizer base + all optimizers. Parity: python/paddle/optimizer/*.py.

State layout follows the reference's accumulator naming so ``.pdopt`` files round-trip:
``{param.name}_{acc}_0`` (moment1, moment2, beta1_pow_acc, velocity, ...), ``master_weights``, ``LR_Scheduler``.

B200 design: for CUDA parameters the update is ONE fused kernel per flat arena (csrc/optim.cu): parameters and
gradients live in contiguous slabs (`enable_flat_arena`), the clip coefficient / loss-scale / found-inf flag are read
from device memory, so an optimizer step never synchronises with the host.
"""
from __future__ import annotations

import math
from collections import OrderedDict, defaultdict

import numpy as np
import torch

from ..framework import dtype as _dt
from ..regularizer import L1Decay, L2Decay
from ..tensor import Parameter, Tensor
from .lr import LRScheduler


def _raw(t):
    return t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) and type(t) is not torch.Tensor else t


class Optimizer:
    _acc_names = ()

    def __init__(self, learning_rate=0.001, parameters=None, weight_decay=None, grad_clip=None, name=None, multi_precision=False):
        if parameters is None:
            from .. import static

            if not static.in_static_mode():
                raise ValueError("parameters must be given in dygraph mode")
            parameters = []   # static mode: collected from the program by minimize()
        parameters = list(parameters)
        self._param_groups = []
        if parameters and isinstance(parameters[0], dict):
            for g in parameters:
                g = dict(g)
                g["params"] = list(g["params"])
                self._param_groups.append(g)
        else:
            self._param_groups.append({"params": parameters})
        self._learning_rate = learning_rate
        self._weight_decay = weight_decay
        self._grad_clip = grad_clip
        self._multi_precision = multi_precision
        self._name = name
        self._accumulators = defaultdict(dict)  # acc_name -> {param.name: tensor}
        self._master_weights = {}
        self._step_count = 0
        self._arena = None
        self._aux = {}
        for g in self._param_groups:
            g.setdefault("learning_rate", 1.0)

    # ---- lr ----------------------------------------------------------------
    def get_lr(self):
        if isinstance(self._learning_rate, LRScheduler):
            return float(self._learning_rate())
        return float(self._learning_rate)

    def set_lr(self, value):
        if isinstance(self._learning_rate, LRScheduler):
            raise RuntimeError("optimizer's learning rate is an LRScheduler; set_lr is not allowed")
        self._learning_rate = float(value)

    def set_lr_scheduler(self, scheduler):
        self._learning_rate = scheduler

    @property
    def _parameter_list(self):
        return [p for g in self._param_groups for p in g["params"]]

    def _wd_value(self, group, p):
        if getattr(p, "regularizer", None) is not None:
            wd = p.regularizer
        else:
            wd = group.get("weight_decay", self._weight_decay)
        if wd is None:
            return 0.0, "l2"
        if isinstance(wd, L1Decay):
            return float(wd.coeff), "l1"
        if isinstance(wd, L2Decay):
            return float(wd.coeff), "l2"
        return float(wd), "l2"

    # ---- state -------------------------------------------------------------
    def _acc(self, name, p, init=0.0, dtype=None, shape=None):
        d = self._accumulators[name]
        t = d.get(p.name)
        if t is None:
            dtype = dtype or (torch.float32 if (self._multi_precision or p.dtype == torch.float32) else p.dtype)
            t = torch.full(tuple(p.size()) if shape is None else shape, init, dtype=dtype, device=p.device)
            d[p.name] = t
        return t

    def _master(self, p):
        if not self._multi_precision or p.dtype == torch.float32:
            return None
        m = self._master_weights.get(p.name)
        if m is None:
            m = _raw(p).detach().float().clone()
            self._master_weights[p.name] = m
        return m

    def state_dict(self):
        sd = OrderedDict()
        byname = {p.name: p for p in self._parameter_list} if self._parameter_list and not isinstance(self._parameter_list[0], dict) else {}

        def tag(pname, t):
            t = t.as_subclass(Tensor)
            p = byname.get(pname)
            ds = p.__dict__.get("_dist_shard") if p is not None else None
            if ds is not None and tuple(t.shape) == tuple(p.shape):     # tensor-parallel parameter: its state is sharded the same way
                t.__dict__["_dist_shard"] = ds
                t.is_distributed = True
            return t

        for acc, d in self._accumulators.items():
            for pname, t in d.items():
                sd[f"{pname}_{acc}_0"] = tag(pname, t)
        if self._master_weights:
            sd["master_weights"] = {k: tag(k, v) for k, v in self._master_weights.items()}
        if isinstance(self._learning_rate, LRScheduler):
            sd["LR_Scheduler"] = self._learning_rate.state_dict()
        sd["@step@"] = self._step_count
        return sd

    def set_state_dict(self, state_dict):
        from ..tensor import _np_to_torch

        def to_t(v, like=None):
            if isinstance(v, tuple) and len(v) == 2 and isinstance(v[1], np.ndarray):
                v = v[1]
            if isinstance(v, np.ndarray):
                v = _np_to_torch(v)
            v = _raw(v)
            return v

        names = {p.name: p for p in self._parameter_list}
        if "LR_Scheduler" in state_dict and isinstance(self._learning_rate, LRScheduler):
            self._learning_rate.set_state_dict(state_dict["LR_Scheduler"])
        if "master_weights" in state_dict:
            for k, v in state_dict["master_weights"].items():
                if k in names:
                    self._master_weights[k] = to_t(v).to(device=names[k].device, dtype=torch.float32).clone()
        self._step_count = int(state_dict.get("@step@", self._step_count))
        for key, v in state_dict.items():
            if key in ("LR_Scheduler", "master_weights", "@step@"):
                continue
            for acc in self._acc_names:
                suffix = f"_{acc}_0"
                if key.endswith(suffix) and key[: -len(suffix)] in names:
                    p = names[key[: -len(suffix)]]
                    t = to_t(v)
                    cur = self._accumulators[acc].get(p.name)
                    if cur is not None:
                        cur.copy_(t.to(device=cur.device, dtype=cur.dtype).reshape(cur.shape))
                    else:
                        self._accumulators[acc][p.name] = t.to(device=p.device).clone()

    set_dict = set_state_dict

    # ---- main loop ---------------------------------------------------------
    def clear_grad(self, set_to_zero=True):
        if self._arena is not None:
            self._arena.zero_grad()
            return
        for p in self._parameter_list:
            p.clear_grad(set_to_zero=False)

    clear_gradients = clear_grad

    def _collect(self):
        out = []
        for g in self._param_groups:
            for p in g["params"]:
                if p.stop_gradient:
                    continue
                gr = torch.Tensor.grad.__get__(p)
                if gr is None:
                    continue
                out.append((g, p, gr))
        return out

    def _has_sparse_grad(self):
        for g in self._param_groups:
            for p in g["params"]:
                gr = torch.Tensor.grad.__get__(p)
                if gr is not None and gr.layout != torch.strided:
                    return True
        return False

    @torch.no_grad()
    def step(self):
        if self._arena is not None and hasattr(self, "_arena_step") and self._arena_ok() and not self._has_sparse_grad():
            self._step_count += 1
            return self._arena_step()
        items = self._collect()
        if not items:
            return
        self._step_count += 1
        # row-sparse gradients (nn.Embedding(sparse=True)) travel as SelectedRows with merged rows
        from ..framework.selected_rows import SelectedRows

        items = [(g, p, SelectedRows.from_sparse_coo(gr).merge() if gr.layout != torch.strided else gr) for g, p, gr in items]
        if self._grad_clip is not None:
            pg = self._grad_clip([(p, gr if isinstance(gr, SelectedRows) else gr.as_subclass(Tensor)) for _, p, gr in items])
            items = [(g, p, ng if isinstance(ng, SelectedRows) else _raw(ng)) for (g, p, _), (_, ng) in zip(items, pg)]
        base_lr = self.get_lr()
        for g, p, gr in items:
            lr = base_lr * g.get("learning_rate", 1.0) * getattr(p, "optimize_attr", {}).get("learning_rate", 1.0)
            wd, kind = self._wd_value(g, p)
            if isinstance(gr, SelectedRows):
                self._update_param_sparse(g, p, _raw(p), gr, lr, wd, kind)
            else:
                self._update_param(g, p, _raw(p), gr, lr, wd, kind)

    def _update_param_sparse(self, group, p, pr, grad, lr, wd, kind):
        """Row-sparse gradient.  Default: the reference's non-lazy semantics = the dense update with zeros in the rows that were not looked
        up (moments decay everywhere).  SGD and lazy Adam / AdamW override this to touch only `grad.rows`."""
        self._update_param(group, p, pr, grad.to_dense().to(pr.dtype), lr, wd, kind)

    def backward(self, loss, startup_program=None, parameters=None, no_grad_set=None, callbacks=None):
        """First half of minimize(): run autograd and return [(param, grad)]. Parity: optimizer.py:Optimizer.backward."""
        loss.backward()
        params = parameters if parameters is not None else [p for g in self._param_groups for p in g["params"]]
        skip = {getattr(v, "name", v) for v in (no_grad_set or ())}
        out = []
        for p in params:
            g = torch.Tensor.grad.__get__(p)
            if g is not None and p.name not in skip and not p.stop_gradient:
                out.append((p, g.as_subclass(Tensor)))
        return out

    def apply_gradients(self, params_grads):
        """Second half of minimize(): clip / regularise / update with the given (param, grad) pairs."""
        for p, g in params_grads:
            if g is not None and torch.Tensor.grad.__get__(p) is not g:
                torch.Tensor.grad.__set__(p, _raw(g))
        self.step()
        return []

    def append_regularization_ops(self, parameters_and_grads, regularization=None):
        """grad += d(regulariser)/d(param) for every pair (L2Decay -> coeff * p, L1Decay -> coeff * sign(p))."""
        reg = regularization if regularization is not None else self._weight_decay
        out = []
        for p, g in parameters_and_grads:
            r = getattr(p, "regularizer", None) or reg
            if g is None or r is None or isinstance(r, (int, float)) and r == 0:
                out.append((p, g))
                continue
            coeff = float(getattr(r, "_coeff", getattr(r, "coeff", r)) if not isinstance(r, (int, float)) else r)
            l1 = type(r).__name__ == "L1Decay"
            out.append((p, (_raw(g) + coeff * (torch.sign(_raw(p)) if l1 else _raw(p))).as_subclass(Tensor)))
        return out

    def get_opti_var_name_list(self):
        return [t.name if hasattr(t, "name") and isinstance(getattr(t, "name", None), str) else f"{acc}_{pname}"
                for acc, d in self._accumulators.items() for pname, t in d.items()]

    def minimize(self, loss, startup_program=None, parameters=None, no_grad_set=None):
        from .. import static

        prog = static._recording[0]
        if prog is not None and (id(loss) in prog._vids):
            # static graph: defer backward+update to Executor.run (python/paddle/optimizer/optimizer.py:minimize appends ops)
            if parameters is not None:
                self._param_groups[0]["params"] = list(parameters)
            elif not self._parameter_list:
                self._param_groups[0]["params"] = [p for p in prog.all_parameters() if p.requires_grad]
            static._minimize_node(self, loss, prog)
            return None, None
        loss.backward()
        self.step()
        return None, None

    def _apply_decay_to_grad(self, pf, gf, wd, kind):
        if wd == 0.0:
            return gf
        return gf + wd * (pf if kind == "l2" else torch.sign(pf))

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        raise NotImplementedError

    # helper: run an update in fp32 (master weights when enabled) and write back
    def _fp32_view(self, p, pr):
        m = self._master(p)
        return (m, True) if m is not None else ((pr if pr.dtype == torch.float32 else pr.float()), pr.dtype != torch.float32)

    def _write_back(self, p, pr, pf, copy_needed):
        if copy_needed:
            pr.copy_(pf.to(pr.dtype))

    # ---- flat arenas (B200 fast path) --------------------------------------
    def enable_flat_arena(self, arena=None):
        """Move params/grads into contiguous slabs so each step is one fused kernel per group."""
        from ..parallel.arena import ParamArena

        if arena is None:
            fn = getattr(self, "_apply_decay_param_fun", None)
            group_fn = (lambda p: 1 if fn(p.name) else 0) if fn is not None else None
            arena = ParamArena(self._parameter_list, group_fn=group_fn)
            if fn is not None:
                for key, slab in arena.slabs.items():
                    slab.decay = bool(key[2])
        self._arena = arena
        return self._arena


def split_master_join(w_bf16, lo):
    """fp32 master weights from (bf16 parameter, int16 residual): bits = (bf16 bits << 16) + lo  (csrc/optim.cu:split_master_join)."""
    hi = w_bf16.contiguous().view(torch.int16).to(torch.int32) << 16
    return (hi + lo.to(torch.int32)).view(torch.float32)


def split_master_split(master_f32):
    """(bf16 parameter = round-to-nearest of the master, int16 residual) from fp32 master weights."""
    w = master_f32.to(torch.bfloat16)
    d = master_f32.contiguous().view(torch.int32) - (w.view(torch.int16).to(torch.int32) << 16)
    d = torch.where(torch.isnan(master_f32), torch.zeros_like(d), d.clamp(-32768, 32767))
    return w, d.to(torch.int16)


class _NativeOps:
    """csrc/optim.cu launchers."""

    def __init__(self):
        from .._build import ext

        self._e = ext()

    def grad_sq_norm(self, g, out, found_inf):
        self._e.grad_sq_norm(g, out, found_inf)

    def adamw_step(self, *a):
        self._e.adamw_step(*a)

    def adamw_step_dyn(self, *a):
        self._e.adamw_step_dyn(*a)


class _TorchOps:
    """Reference implementation of the same slab ops in plain PyTorch (CPU tests / fallback)."""

    def grad_sq_norm(self, g, out, found_inf):
        s = g.float().pow(2).sum()
        out.add_(s)
        if found_inf is not None and not bool(torch.isfinite(s)):
            found_inf.fill_(1.0)

    def adamw_step(self, p, g, master, m, v, lr, b1, b2, eps, wd, step, sq, max_norm, found_inf, inv_scale):
        if found_inf is not None and float(found_inf) != 0.0:
            return
        gs = float(inv_scale) if inv_scale is not None else 1.0
        if sq is not None and max_norm > 0:
            norm = float(sq.sqrt()) * gs
            if norm > max_norm:
                gs *= max_norm / (norm + 1e-6)
        gf = g.float() * gs
        split = master is not None and master.dtype == torch.int16
        pf = split_master_join(p, master) if split else (master if master is not None else p.float())
        mf, vf = m.float(), v.float()
        mf.mul_(b1).add_(gf, alpha=1 - b1)
        vf.mul_(b2).addcmul_(gf, gf, value=1 - b2)
        c1, c2 = 1 - b1 ** step, 1 - b2 ** step
        pf.mul_(1 - lr * wd).addcdiv_(mf, (vf / c2).sqrt_().add_(eps), value=-lr / c1)
        m.copy_(mf)
        v.copy_(vf)
        if split:
            w, lo = split_master_split(pf)
            p.copy_(w)
            master.copy_(lo)
        else:
            p.copy_(pf)


class SGD(Optimizer):
    def __init__(self, learning_rate=0.001, parameters=None, weight_decay=None, grad_clip=None, multi_precision=False, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name, multi_precision)

    def _update_param_sparse(self, group, p, pr, grad, lr, wd, kind):
        # paddle/phi/kernels/selected_rows/sgd_kernel: param[rows] -= lr * grad_rows (regularisation reaches only the touched rows too)
        if wd != 0.0 or pr.dtype != grad.value.dtype or self._master(p) is not None:
            return super()._update_param_sparse(group, p, pr, grad, lr, wd, kind)
        pr.index_add_(0, grad.rows.to(pr.device), grad.value, alpha=-lr)

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        if pr.is_cuda and kind == "l2" and pr.dtype in (torch.float32, torch.bfloat16, torch.float16) and grad.dtype == pr.dtype:
            from .._build import ext

            ext().sgd_step(pr, grad.contiguous(), self._master(p), None, lr, 0.0, wd, False)
            return
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float(), wd, kind)
        pf.add_(gf, alpha=-lr)
        self._write_back(p, pr, pf, cp)


class Momentum(Optimizer):
    _acc_names = ("velocity",)

    def __init__(self, learning_rate=0.001, momentum=0.9, parameters=None, use_nesterov=False, weight_decay=None, grad_clip=None,
                 multi_precision=False, rescale_grad=1.0, use_multi_tensor=False, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name, multi_precision)
        self._momentum, self._use_nesterov, self._rescale_grad = momentum, use_nesterov, rescale_grad

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        v = self._acc("velocity", p, dtype=torch.float32)
        if pr.is_cuda and kind == "l2" and self._rescale_grad == 1.0 and pr.dtype in (torch.float32, torch.bfloat16, torch.float16) and grad.dtype == pr.dtype:
            from .._build import ext

            ext().sgd_step(pr, grad.contiguous(), self._master(p), v, lr, self._momentum, wd, self._use_nesterov)
            return
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float() * self._rescale_grad, wd, kind)
        v.mul_(self._momentum).add_(gf)
        if self._use_nesterov:
            pf.add_(gf + self._momentum * v, alpha=-lr)
        else:
            pf.add_(v, alpha=-lr)
        self._write_back(p, pr, pf, cp)


class Adam(Optimizer):
    _acc_names = ("moment1", "moment2", "beta1_pow_acc", "beta2_pow_acc", "moment2_max")
    _decoupled = False

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-08, parameters=None, weight_decay=None, grad_clip=None,
                 lazy_mode=False, multi_precision=False, use_multi_tensor=False, amsgrad=False, name=None, moment_dtype=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name, multi_precision)
        self._beta1, self._beta2, self._epsilon, self._amsgrad = beta1, beta2, epsilon, amsgrad
        self._lazy_mode = bool(lazy_mode)     # row-sparse gradients update only their rows (moments of the other rows stay frozen)
        self._moment_dtype = _dt.convert_dtype(moment_dtype)
        self._lr_ratio = None
        self._apply_decay_param_fun = None

    def _betas(self):
        b1 = self._beta1.item() if isinstance(self._beta1, torch.Tensor) else self._beta1
        b2 = self._beta2.item() if isinstance(self._beta2, torch.Tensor) else self._beta2
        return float(b1), float(b2)

    def _state_dtype(self, p):
        if self._moment_dtype is not None:
            return self._moment_dtype
        return torch.float32 if (self._multi_precision or p.dtype == torch.float32) else p.dtype

    def _update_param_sparse(self, group, p, pr, grad, lr, wd, kind):
        """lazy_mode=True (paddle/phi/kernels/selected_rows/adam_kernel, lazy branch): only the rows present in the gradient update their
        moments and their parameter rows; the bias correction still follows the global step.  Otherwise: dense semantics."""
        if not getattr(self, "_lazy_mode", False) or self._amsgrad or self._master(p) is not None:
            return super()._update_param_sparse(group, p, pr, grad, lr, wd, kind)
        b1, b2 = self._betas()
        sdt = self._state_dtype(p)
        m = self._acc("moment1", p, dtype=sdt)
        v = self._acc("moment2", p, dtype=sdt)
        b1p = self._acc("beta1_pow_acc", p, init=1.0, dtype=torch.float32, shape=(1,))
        b2p = self._acc("beta2_pow_acc", p, init=1.0, dtype=torch.float32, shape=(1,))
        step = self._aux.setdefault("steps", {}).get(p.name, 0) + 1
        self._aux["steps"][p.name] = step
        rows = grad.rows.to(pr.device)
        gf = grad.value.float()
        pf = pr[rows].float()
        if self._decoupled:
            if self._apply_decay_param_fun is not None and not self._apply_decay_param_fun(p.name):
                wd = 0.0
            pf = pf * (1.0 - lr * wd)
        elif wd != 0.0:
            gf = self._apply_decay_to_grad(pf, gf, wd, kind)
        mf = m[rows].float() * b1 + gf * (1 - b1)
        vf = v[rows].float() * b2 + gf * gf * (1 - b2)
        b1p.mul_(b1)
        b2p.mul_(b2)
        c1, c2 = 1 - b1 ** step, 1 - b2 ** step
        pf = pf - (lr / c1) * mf / ((vf / c2).sqrt() + self._epsilon)
        m[rows] = mf.to(m.dtype)
        v[rows] = vf.to(v.dtype)
        pr[rows] = pf.to(pr.dtype)

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        b1, b2 = self._betas()
        sdt = self._state_dtype(p)
        m = self._acc("moment1", p, dtype=sdt)
        v = self._acc("moment2", p, dtype=sdt)
        b1p = self._acc("beta1_pow_acc", p, init=1.0, dtype=torch.float32, shape=(1,))
        b2p = self._acc("beta2_pow_acc", p, init=1.0, dtype=torch.float32, shape=(1,))
        step = self._aux.setdefault("steps", {}).get(p.name, 0) + 1
        self._aux["steps"][p.name] = step
        if self._decoupled:
            if self._apply_decay_param_fun is not None and not self._apply_decay_param_fun(p.name):
                wd = 0.0
            if self._lr_ratio is not None:
                lr = lr * self._lr_ratio(p)
        fused = (pr.is_cuda and not self._amsgrad and pr.dtype in (torch.float32, torch.bfloat16, torch.float16)
                 and grad.dtype in (pr.dtype, torch.float32) and sdt in (torch.float32, torch.bfloat16) and pr.is_contiguous()
                 and (self._decoupled or wd == 0.0))
        if fused:
            from .._build import ext

            ext().adamw_step(pr, grad.contiguous(), self._master(p), m, v, lr, b1, b2, float(self._epsilon),
                             wd if self._decoupled else 0.0, step, None, 0.0, None, None)
            b1p.mul_(b1)
            b2p.mul_(b2)
            return
        pf, cp = self._fp32_view(p, pr)
        gf = grad.float()
        if not self._decoupled:
            gf = self._apply_decay_to_grad(pf, gf, wd, kind)
        else:
            pf.mul_(1.0 - lr * wd)
        mf, vf = (m if m.dtype == torch.float32 else m.float()), (v if v.dtype == torch.float32 else v.float())
        mf.mul_(b1).add_(gf, alpha=1 - b1)
        vf.mul_(b2).addcmul_(gf, gf, value=1 - b2)
        b1p.mul_(b1)
        b2p.mul_(b2)
        c1, c2 = 1 - b1 ** step, 1 - b2 ** step
        if self._amsgrad:
            vmax = self._acc("moment2_max", p, dtype=torch.float32)
            torch.maximum(vmax, vf, out=vmax)
            denom = (vmax / c2).sqrt_().add_(self._epsilon)
        else:
            denom = (vf / c2).sqrt_().add_(self._epsilon)
        pf.addcdiv_(mf, denom, value=-lr / c1)
        if m.dtype != torch.float32:
            m.copy_(mf)
            v.copy_(vf)
        self._write_back(p, pr, pf, cp)

    # ---- flat-arena fast path: one fused kernel per slab, clip/loss-scale read on device ---------------------
    def _arena_ok_static(self):
        from ..nn.clip import ClipGradByGlobalNorm

        if self._amsgrad or not (self._decoupled or not self._weight_decay):
            return False
        return self._grad_clip is None or isinstance(self._grad_clip, ClipGradByGlobalNorm)

    def _arena_ok(self):
        from ..nn.clip import ClipGradByGlobalNorm

        if self._amsgrad or not (self._decoupled or not self._weight_decay):
            return False
        if self._grad_clip is not None and not isinstance(self._grad_clip, ClipGradByGlobalNorm):
            return False
        if getattr(self, "_lr_ratio", None) is not None or len(self._param_groups) > 1:
            return False      # per-parameter lr scaling / param-group overrides: the one-launch-per-slab update cannot express them
        ok = self._aux.get("arena_params_ok")
        if ok is None:        # per-parameter attributes the slab kernel ignores (checked once: they do not change after construction)
            ok = True
            for sl in self._arena.all_slabs():
                for p in sl.params:
                    if getattr(p, "optimize_attr", {}).get("learning_rate", 1.0) != 1.0 or not getattr(p, "need_clip", True):
                        ok = False
            self._aux["arena_params_ok"] = ok
        return ok

    def _shard_bounds(self, numel):
        """[lo, hi) of a slab that this rank updates under optimizer-state sharding (stage 1); the whole slab otherwise."""
        sh = self._aux.get("shard")
        if sh is None:
            return 0, numel
        rank, world, _ = sh
        per = ((numel + world - 1) // world + 7) // 8 * 8          # 16-byte aligned shard starts for the vectorised kernel
        return min(rank * per, numel), min((rank + 1) * per, numel)

    def _arena_step(self):
        cuda = all(s.data.is_cuda for s in self._arena.all_slabs())
        E = _NativeOps() if cuda else _TorchOps()   # same slab algorithm; CPU path = plain torch (used by the gloo tests)
        b1, b2 = self._betas()
        lr = self.get_lr()
        slabs = self._arena.all_slabs()
        dev = slabs[0].data.device
        shard = self._aux.get("shard")     # (rank, world, group): sharding stage 1 — every rank owns 1/world of each slab's state
        pending = self._aux.pop("pending_state", None)
        if pending is not None:
            self._arena_restore(pending)
        sq = None
        max_norm = 0.0
        if self._grad_clip is not None:
            sq = self._aux.get("sq")
            if sq is None:
                sq = self._aux["sq"] = torch.zeros(1, dtype=torch.float32, device=dev)
                self._aux["found_inf"] = torch.zeros(1, dtype=torch.float32, device=dev)
            sq.zero_()
            for s in slabs:
                if getattr(s, "clip_weight", 1.0) != 0.0:
                    lo, hi = self._shard_bounds(s.numel)      # sharded: own range only, the hook sums over the sharding group
                    if hi > lo:
                        E.grad_sq_norm(s.grad[lo:hi], sq, self._aux["found_inf"])
            hook = self._aux.get("norm_allreduce")
            if hook is not None:
                hook(sq)          # hybrid parallel: sum the squared norm over mp/pp/sharding groups
            max_norm = float(self._grad_clip.clip_norm)
        found_inf = self._aux.get("scaler_found_inf")
        inv_scale = self._aux.get("scaler_inv_scale")
        for s in slabs:
            lo, hi = self._shard_bounds(s.numel)
            n = hi - lo
            if self._multi_precision and s.dtype != torch.float32 and s.master is None:
                s.master = self._new_master(s, lo, hi)
            if "m" not in s.state:
                sdt = self._moment_dtype or (torch.float32 if (self._multi_precision or s.dtype == torch.float32) else s.dtype)
                s.state["m"] = torch.zeros(n, dtype=sdt, device=dev)
                s.state["v"] = torch.zeros(n, dtype=sdt, device=dev)
            if n > 0:
                wd = float(self._weight_decay or 0.0) if getattr(s, "decay", True) else 0.0
                data, grad = (s.data, s.grad) if shard is None else (s.data[lo:hi], s.grad[lo:hi])
                hp = self._aux.get("dyn_hparams") if cuda else None
                if hp is not None:   # graph-capturable launch: lr and bias corrections are read from device memory (jit.CapturedTrainStep)
                    E.adamw_step_dyn(data, grad, s.master, s.state["m"], s.state["v"], float(getattr(s, "lr_scale", 1.0)), b1, b2,
                                     float(self._epsilon), wd, sq, max_norm, found_inf, inv_scale, hp)
                else:
                    E.adamw_step(data, grad, s.master, s.state["m"], s.state["v"], lr * getattr(s, "lr_scale", 1.0), b1, b2,
                                 float(self._epsilon), wd, self._step_count, sq, max_norm, found_inf, inv_scale)
            if shard is not None:
                self._gather_shards(s, shard)

    def _arena_restore(self, state_dict):
        """Checkpoint -> slab state (moments, fp32 master weights), honouring the shard range. Called from set_state_dict; when the
        arena is created later the dict waits in `_aux['pending_state']` until the first step."""
        slabs = self._arena.all_slabs()
        dev = slabs[0].data.device
        masters = state_dict.get("master_weights", {}) or {}
        for s in slabs:
            lo, hi = self._shard_bounds(s.numel)
            n = hi - lo
            if not any(f"{p.name}_moment1_0" in state_dict for p in s.params):
                continue
            sdt = self._moment_dtype or (torch.float32 if (self._multi_precision or s.dtype == torch.float32) else s.dtype)
            if "m" not in s.state:
                s.state["m"] = torch.zeros(n, dtype=sdt, device=dev)
                s.state["v"] = torch.zeros(n, dtype=sdt, device=dev)
            if self._multi_precision and s.dtype != torch.float32 and s.master is None:
                s.master = self._new_master(s, lo, hi)
            for p in s.params:
                o, cnt = s.offsets[p.name]
                a, b = max(o, lo), min(o + cnt, hi)          # part of this parameter that falls into the local range
                if a >= b:
                    continue
                for key, dst in ((f"{p.name}_moment1_0", s.state["m"]), (f"{p.name}_moment2_0", s.state["v"])):
                    src = state_dict.get(key)
                    if src is not None:
                        dst[a - lo:b - lo].copy_(torch.as_tensor(np.asarray(src) if not isinstance(src, torch.Tensor) else src).reshape(-1)[a - o:b - o].to(dst.dtype))
                mw = masters.get(p.name)
                if mw is not None and s.master is not None:
                    mwf = torch.as_tensor(np.asarray(mw) if not isinstance(mw, torch.Tensor) else mw).reshape(-1)[a - o:b - o].float().to(dev)
                    if s.master.dtype == torch.int16:      # split master: the parameter slab carries the high half
                        w, r = split_master_split(mwf)
                        s.data[a:b].copy_(w)
                        s.master[a - lo:b - lo].copy_(r)
                    else:
                        s.master[a - lo:b - lo].copy_(mwf)

    def _new_master(self, s, lo, hi):
        """fp32 master weights of the slab range [lo, hi).  bf16 slabs use the split format (int16 residual next to the bf16 parameter:
        4 instead of 6 bytes per parameter, see csrc/optim.cu) unless FLAGS_b200_split_master_weights is off."""
        from ..framework.flags import flag

        if s.dtype == torch.bfloat16 and flag("FLAGS_b200_split_master_weights", True):
            return torch.zeros(hi - lo, dtype=torch.int16, device=s.data.device)      # residual 0: the master starts equal to the parameter
        return s.data[lo:hi].float()

    def _gather_shards(self, slab, shard):
        """Every owner publishes its updated range of the parameter slab (stage 1: parameters stay replicated)."""
        import torch.distributed as dist

        rank, world, group = shard
        pg = getattr(group, "pg", group)
        saved = self._aux["shard"]
        for r in range(world):
            self._aux["shard"] = (r, world, group)
            lo, hi = self._shard_bounds(slab.numel)
            if hi > lo:
                dist.broadcast(slab.data[lo:hi], src=dist.get_global_rank(pg, r) if pg is not None else r, group=pg)
        self._aux["shard"] = saved

    def _full_state(self, slab, key):
        """Full-length view of a (possibly sharded) per-slab state tensor: shards are gathered for checkpoints."""
        shard = self._aux.get("shard")
        t = slab.master if key == "master" else slab.state[key]
        if key == "master" and t is not None and t.dtype == torch.int16:
            lo, hi = self._shard_bounds(slab.numel)
            t = split_master_join(slab.data[lo:hi], t)        # checkpoints always carry fp32 master weights
        if shard is None or t is None:
            return t
        import torch.distributed as dist

        rank, world, group = shard
        pg = getattr(group, "pg", group)
        full = torch.zeros(slab.numel, dtype=t.dtype, device=t.device)
        saved = self._aux["shard"]
        for r in range(world):
            self._aux["shard"] = (r, world, group)
            lo, hi = self._shard_bounds(slab.numel)
            if hi > lo:
                if r == rank:
                    full[lo:hi].copy_(t)
                dist.broadcast(full[lo:hi], src=dist.get_global_rank(pg, r) if pg is not None else r, group=pg)
        self._aux["shard"] = saved
        return full

    def _refresh_dyn_hparams(self):
        """Write {lr, 1 - b1^t, 1 - b2^t} for the *next* update into the device tensor the captured AdamW launches read."""
        hp = self._aux.get("dyn_hparams")
        if hp is None:
            dev = self._arena.all_slabs()[0].data.device
            hp = self._aux["dyn_hparams"] = torch.zeros(3, dtype=torch.float32, device=dev)
            self._aux["dyn_hparams_host"] = torch.zeros(3, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(3)
        b1, b2 = self._betas()
        t = self._step_count + 1
        host = self._aux["dyn_hparams_host"]
        host[0], host[1], host[2] = float(self.get_lr()), 1.0 - b1 ** t, 1.0 - b2 ** t
        hp.copy_(host, non_blocking=True)
        return hp

    def state_dict(self):
        sd = super().state_dict()
        if self._arena is not None:
            for s in self._arena.all_slabs():
                if "m" not in s.state:
                    continue
                fm, fv, fmaster = self._full_state(s, "m"), self._full_state(s, "v"), self._full_state(s, "master")
                for p in s.params:
                    o, n = s.offsets[p.name]
                    sd[f"{p.name}_moment1_0"] = fm[o:o + n].view(tuple(p.size())).as_subclass(Tensor)
                    sd[f"{p.name}_moment2_0"] = fv[o:o + n].view(tuple(p.size())).as_subclass(Tensor)
                    if fmaster is not None:
                        sd.setdefault("master_weights", {})[p.name] = fmaster[o:o + n].view(tuple(p.size())).as_subclass(Tensor)
                    b1, b2 = self._betas()
                    sd[f"{p.name}_beta1_pow_acc_0"] = torch.tensor([b1 ** self._step_count], dtype=torch.float32).as_subclass(Tensor)
                    sd[f"{p.name}_beta2_pow_acc_0"] = torch.tensor([b2 ** self._step_count], dtype=torch.float32).as_subclass(Tensor)
                    ds = p.__dict__.get("_dist_shard")
                    if ds is not None:       # tensor-parallel parameter: its state is sharded the same way (distributed/checkpoint.py)
                        for t in (sd[f"{p.name}_moment1_0"], sd[f"{p.name}_moment2_0"], sd.get("master_weights", {}).get(p.name)):
                            if t is not None:
                                t.__dict__["_dist_shard"] = ds
                                t.is_distributed = True
        return sd

    def set_state_dict(self, state_dict):
        super().set_state_dict(state_dict)
        if self._arena is not None:
            with torch.no_grad():
                self._arena_restore(state_dict)
        else:
            self._aux["pending_state"] = state_dict      # the flat arena may be enabled after loading
        steps = self._aux.setdefault("steps", {})
        b1, _ = self._betas()
        derived = 0
        for pname, t in self._accumulators.get("beta1_pow_acc", {}).items():
            val = float(t.reshape(-1)[0])
            if 0 < val < 1 and 0 < b1 < 1:
                steps[pname] = max(steps.get(pname, 0), int(round(math.log(val) / math.log(b1))))
                derived = max(derived, steps[pname])
        if "@step@" not in state_dict and derived:
            self._step_count = derived          # a reference .pdopt has no '@step@': the bias correction continues from beta1_pow_acc
        elif "@step@" in state_dict:
            for p in self._parameter_list if self._parameter_list and not isinstance(self._parameter_list[0], dict) else []:
                steps.setdefault(p.name, int(self._step_count))   # arena-written checkpoints loaded on the per-parameter path


class AdamW(Adam):
    _decoupled = True

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-08, parameters=None, weight_decay=0.01, lr_ratio=None,
                 apply_decay_param_fun=None, grad_clip=None, lazy_mode=False, multi_precision=False, amsgrad=False, name=None, moment_dtype=None):
        super().__init__(learning_rate, beta1, beta2, epsilon, parameters, weight_decay, grad_clip, lazy_mode, multi_precision, False, amsgrad, name, moment_dtype)
        self._lr_ratio, self._apply_decay_param_fun = lr_ratio, apply_decay_param_fun


class Adamax(Optimizer):
    _acc_names = ("moment", "inf_norm", "beta1_pow_acc")

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-08, parameters=None, weight_decay=None, grad_clip=None, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name)
        self._beta1, self._beta2, self._epsilon = beta1, beta2, epsilon

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        m, u = self._acc("moment", p, dtype=torch.float32), self._acc("inf_norm", p, dtype=torch.float32)
        b1p = self._acc("beta1_pow_acc", p, init=1.0, dtype=torch.float32, shape=(1,))
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float(), wd, kind)
        m.mul_(self._beta1).add_(gf, alpha=1 - self._beta1)
        torch.maximum(u * self._beta2, gf.abs() + self._epsilon, out=u)
        b1p.mul_(self._beta1)
        pf.addcdiv_(m, u, value=-lr / (1 - float(b1p)))
        self._write_back(p, pr, pf, cp)


class Adagrad(Optimizer):
    _acc_names = ("moment",)

    def __init__(self, learning_rate, epsilon=1e-06, parameters=None, weight_decay=None, grad_clip=None, name=None, initial_accumulator_value=0.0):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name)
        self._epsilon, self._init = epsilon, initial_accumulator_value

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        m = self._acc("moment", p, init=self._init, dtype=torch.float32)
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float(), wd, kind)
        m.addcmul_(gf, gf)
        pf.addcdiv_(gf, m.sqrt() + self._epsilon, value=-lr)
        self._write_back(p, pr, pf, cp)


class Adadelta(Optimizer):
    _acc_names = ("_avg_squared_grad", "_avg_squared_update")

    def __init__(self, learning_rate=0.001, epsilon=1e-06, rho=0.95, parameters=None, weight_decay=None, grad_clip=None, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name)
        self._epsilon, self._rho = epsilon, rho

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        eg, ed = self._acc("_avg_squared_grad", p, dtype=torch.float32), self._acc("_avg_squared_update", p, dtype=torch.float32)
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float(), wd, kind)
        eg.mul_(self._rho).addcmul_(gf, gf, value=1 - self._rho)
        upd = gf * ((ed + self._epsilon).sqrt() / (eg + self._epsilon).sqrt())
        ed.mul_(self._rho).addcmul_(upd, upd, value=1 - self._rho)
        pf.add_(upd, alpha=-lr)
        self._write_back(p, pr, pf, cp)


class RMSProp(Optimizer):
    _acc_names = ("momentum", "mean_square", "mean_grad")

    def __init__(self, learning_rate, rho=0.95, epsilon=1e-06, momentum=0.0, centered=False, parameters=None, weight_decay=None, grad_clip=None, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name)
        self._rho, self._epsilon, self._momentum, self._centered = rho, epsilon, momentum, centered

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        ms, mom = self._acc("mean_square", p, dtype=torch.float32), self._acc("momentum", p, dtype=torch.float32)
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float(), wd, kind)
        ms.mul_(self._rho).addcmul_(gf, gf, value=1 - self._rho)
        if self._centered:
            mg = self._acc("mean_grad", p, dtype=torch.float32)
            mg.mul_(self._rho).add_(gf, alpha=1 - self._rho)
            denom = (ms - mg * mg + self._epsilon).sqrt()
        else:
            denom = (ms + self._epsilon).sqrt()
        mom.mul_(self._momentum).addcdiv_(gf, denom, value=lr)
        pf.sub_(mom)
        self._write_back(p, pr, pf, cp)


class Lamb(Optimizer):
    _acc_names = ("moment1", "moment2", "beta1_pow_acc", "beta2_pow_acc")

    def __init__(self, learning_rate=0.001, lamb_weight_decay=0.01, beta1=0.9, beta2=0.999, epsilon=1e-06, parameters=None, grad_clip=None,
                 exclude_from_weight_decay_fn=None, multi_precision=False, always_adapt=False, name=None):
        super().__init__(learning_rate, parameters, None, grad_clip, name, multi_precision)
        self._lamb_wd, self._beta1, self._beta2, self._epsilon = lamb_weight_decay, beta1, beta2, epsilon
        self._exclude_fn, self._always_adapt = exclude_from_weight_decay_fn, always_adapt

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        m, v = self._acc("moment1", p, dtype=torch.float32), self._acc("moment2", p, dtype=torch.float32)
        step = self._aux.setdefault("steps", {}).get(p.name, 0) + 1
        self._aux["steps"][p.name] = step
        wd = 0.0 if (self._exclude_fn is not None and self._exclude_fn(p)) else self._lamb_wd
        if pr.is_cuda and pr.dtype in (torch.float32, torch.bfloat16, torch.float16) and grad.dtype == pr.dtype and pr.is_contiguous():
            from .._build import ext

            ext().lamb_step(pr, grad.contiguous(), self._master(p), m, v, lr, self._beta1, self._beta2, self._epsilon, wd, step)
            return
        pf, cp = self._fp32_view(p, pr)
        gf = grad.float()
        m.mul_(self._beta1).add_(gf, alpha=1 - self._beta1)
        v.mul_(self._beta2).addcmul_(gf, gf, value=1 - self._beta2)
        mh, vh = m / (1 - self._beta1 ** step), v / (1 - self._beta2 ** step)
        r = mh / (vh.sqrt() + self._epsilon) + wd * pf
        pn, rn = pf.norm(), r.norm()
        trust = torch.where((pn > 0) & (rn > 0), pn / rn, torch.ones_like(pn))
        pf.add_(r * trust, alpha=-lr)
        self._write_back(p, pr, pf, cp)


class NAdam(Optimizer):
    _acc_names = ("moment1", "moment2", "mu_product")

    def __init__(self, learning_rate=0.002, beta1=0.9, beta2=0.999, epsilon=1e-08, momentum_decay=0.004, parameters=None, weight_decay=None, grad_clip=None, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name)
        self._beta1, self._beta2, self._epsilon, self._md = beta1, beta2, epsilon, momentum_decay

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        m, v = self._acc("moment1", p, dtype=torch.float32), self._acc("moment2", p, dtype=torch.float32)
        mup = self._acc("mu_product", p, init=1.0, dtype=torch.float32, shape=(1,))
        step = self._aux.setdefault("steps", {}).get(p.name, 0) + 1
        self._aux["steps"][p.name] = step
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float(), wd, kind)
        mu = self._beta1 * (1 - 0.5 * 0.96 ** (step * self._md))
        mu_next = self._beta1 * (1 - 0.5 * 0.96 ** ((step + 1) * self._md))
        mup.mul_(mu)
        mp = float(mup)
        m.mul_(self._beta1).add_(gf, alpha=1 - self._beta1)
        v.mul_(self._beta2).addcmul_(gf, gf, value=1 - self._beta2)
        denom = (v / (1 - self._beta2 ** step)).sqrt() + self._epsilon
        pf.addcdiv_(gf, denom, value=-lr * (1 - mu) / (1 - mp))
        pf.addcdiv_(m, denom, value=-lr * mu_next / (1 - mp * mu_next))
        self._write_back(p, pr, pf, cp)


class RAdam(Optimizer):
    _acc_names = ("moment1", "moment2")

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-08, parameters=None, weight_decay=None, grad_clip=None, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name)
        self._beta1, self._beta2, self._epsilon = beta1, beta2, epsilon

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        m, v = self._acc("moment1", p, dtype=torch.float32), self._acc("moment2", p, dtype=torch.float32)
        step = self._aux.setdefault("steps", {}).get(p.name, 0) + 1
        self._aux["steps"][p.name] = step
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float(), wd, kind)
        m.mul_(self._beta1).add_(gf, alpha=1 - self._beta1)
        v.mul_(self._beta2).addcmul_(gf, gf, value=1 - self._beta2)
        mh = m / (1 - self._beta1 ** step)
        rho_inf = 2 / (1 - self._beta2) - 1
        rho_t = rho_inf - 2 * step * self._beta2 ** step / (1 - self._beta2 ** step)
        if rho_t > 5:
            l = math.sqrt(1 - self._beta2 ** step) / (v.sqrt() + self._epsilon)
            r = math.sqrt((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t))
            pf.add_(mh * l * r, alpha=-lr)
        else:
            pf.add_(mh, alpha=-lr)
        self._write_back(p, pr, pf, cp)


class ASGD(Optimizer):
    _acc_names = ("d", "y")

    def __init__(self, learning_rate=0.001, batch_num=1, parameters=None, weight_decay=None, grad_clip=None, multi_precision=False, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name, multi_precision)
        self._batch_num = batch_num

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        d = self._acc("d", p, dtype=torch.float32)
        ys = self._acc("y", p, dtype=torch.float32, shape=(self._batch_num, *p.size()))
        idx = (self._step_count - 1) % self._batch_num
        pf, cp = self._fp32_view(p, pr)
        gf = self._apply_decay_to_grad(pf, grad.float(), wd, kind)
        d.add_(gf - ys[idx])
        ys[idx].copy_(gf)
        n = min(self._step_count, self._batch_num)
        pf.add_(d, alpha=-lr / n)
        self._write_back(p, pr, pf, cp)


class Rprop(Optimizer):
    _acc_names = ("prev_grad", "step_size")

    def __init__(self, learning_rate=0.001, learning_rate_range=(1e-5, 50), parameters=None, etas=(0.5, 1.2), grad_clip=None, multi_precision=False, name=None):
        super().__init__(learning_rate, parameters, None, grad_clip, name, multi_precision)
        self._range, self._etas = learning_rate_range, etas

    def _update_param(self, group, p, pr, grad, lr, wd, kind):
        prev = self._acc("prev_grad", p, dtype=torch.float32)
        ss = self._acc("step_size", p, init=lr, dtype=torch.float32)
        pf, cp = self._fp32_view(p, pr)
        gf = grad.float().clone()
        sign = (gf * prev).sign()
        ss.mul_(torch.where(sign > 0, self._etas[1], torch.where(sign < 0, self._etas[0], 1.0))).clamp_(self._range[0], self._range[1])
        gf[sign < 0] = 0
        pf.addcmul_(gf.sign(), ss, value=-1)
        prev.copy_(gf)
        self._write_back(p, pr, pf, cp)


class LBFGS(Optimizer):
    """Parity: python/paddle/optimizer/lbfgs.py (closure-based, strong-Wolfe optional)."""

    def __init__(self, learning_rate=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-07, tolerance_change=1e-09, history_size=100,
                 line_search_fn=None, parameters=None, weight_decay=None, grad_clip=None, name=None):
        super().__init__(learning_rate, parameters, weight_decay, grad_clip, name)
        ps = [torch.nn.Parameter(_raw(p).detach()) for p in self._parameter_list]
        self._shadow = ps
        self._inner = torch.optim.LBFGS(ps, lr=float(learning_rate), max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                                        tolerance_change=tolerance_change, history_size=history_size, line_search_fn=line_search_fn)

    def step(self, closure):
        params = self._parameter_list

        def inner_closure():
            with torch.no_grad():
                for p, s in zip(params, self._shadow):
                    torch.Tensor.copy_(p, s)
            for p in params:
                p.clear_grad()
            with torch.enable_grad():
                loss = closure()
            for p, s in zip(params, self._shadow):
                g = torch.Tensor.grad.__get__(p)
                s.grad = None if g is None else g.detach().clone()
            return _raw(loss).detach()

        loss = self._inner.step(inner_closure)
        with torch.no_grad():
            for p, s in zip(params, self._shadow):
                torch.Tensor.copy_(p, s)
        return loss.as_subclass(Tensor) if isinstance(loss, torch.Tensor) else loss
