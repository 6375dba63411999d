"""paddle.metric. Parity: python/paddle/metric/metrics.py."""
from __future__ import annotations

import numpy as np
import torch

from ..tensor import Tensor


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().float().cpu().as_subclass(torch.Tensor).numpy()
    return np.asarray(x)


class Metric:
    def reset(self):
        raise NotImplementedError

    def update(self, *args):
        raise NotImplementedError

    def accumulate(self):
        raise NotImplementedError

    def name(self):
        raise NotImplementedError

    def compute(self, *args):
        return args


class Accuracy(Metric):
    def __init__(self, topk=(1,), name=None, *args, **kwargs):
        self.topk = topk if isinstance(topk, (list, tuple)) else (topk,)
        self.maxk = max(self.topk)
        self._name = name or "acc"
        self.reset()

    def compute(self, pred, label, *args):
        pred = pred.as_subclass(torch.Tensor) if isinstance(pred, torch.Tensor) else torch.as_tensor(pred)
        label = label.as_subclass(torch.Tensor) if isinstance(label, torch.Tensor) else torch.as_tensor(label)
        idx = pred.topk(self.maxk, -1).indices
        if label.dim() == pred.dim() and label.shape[-1] != 1:
            label = label.argmax(-1, keepdim=True)
        label = label.reshape(-1, 1)
        return (idx == label).float().as_subclass(Tensor)

    def update(self, correct, *args):
        c = _np(correct)
        accs = []
        for i, k in enumerate(self.topk):
            num = c[..., :k].sum()
            self.total[i] += num
            self.count[i] += c.shape[0]
            accs.append(float(num) / max(1, c.shape[0]))
        return accs[0] if len(accs) == 1 else accs

    def reset(self):
        self.total = [0.0] * len(self.topk)
        self.count = [0] * len(self.topk)

    def accumulate(self):
        r = [t / c if c > 0 else 0.0 for t, c in zip(self.total, self.count)]
        return r[0] if len(r) == 1 else r

    def name(self):
        return [f"{self._name}_top{k}" for k in self.topk] if len(self.topk) > 1 else [self._name]


class Precision(Metric):
    def __init__(self, name="precision", *args, **kwargs):
        self._name = name
        self.reset()

    def update(self, preds, labels):
        p = (_np(preds).reshape(-1) > 0.5).astype(np.int64)
        l = _np(labels).reshape(-1).astype(np.int64)
        self.tp += int(((p == 1) & (l == 1)).sum())
        self.fp += int(((p == 1) & (l == 0)).sum())

    def reset(self):
        self.tp = self.fp = 0

    def accumulate(self):
        return self.tp / (self.tp + self.fp) if self.tp + self.fp else 0.0

    def name(self):
        return self._name


class Recall(Metric):
    def __init__(self, name="recall", *args, **kwargs):
        self._name = name
        self.reset()

    def update(self, preds, labels):
        p = (_np(preds).reshape(-1) > 0.5).astype(np.int64)
        l = _np(labels).reshape(-1).astype(np.int64)
        self.tp += int(((p == 1) & (l == 1)).sum())
        self.fn += int(((p == 0) & (l == 1)).sum())

    def reset(self):
        self.tp = self.fn = 0

    def accumulate(self):
        return self.tp / (self.tp + self.fn) if self.tp + self.fn else 0.0

    def name(self):
        return self._name


class Auc(Metric):
    def __init__(self, curve="ROC", num_thresholds=4095, name="auc", *args, **kwargs):
        self._curve, self._n, self._name = curve, num_thresholds, name
        self.reset()

    def update(self, preds, labels):
        p = _np(preds)
        p = p[:, 1] if p.ndim == 2 and p.shape[1] == 2 else p.reshape(-1)
        l = _np(labels).reshape(-1)
        bins = np.clip((p * self._n).astype(np.int64), 0, self._n)
        np.add.at(self._pos, bins[l > 0.5], 1)
        np.add.at(self._neg, bins[l <= 0.5], 1)

    def reset(self):
        self._pos = np.zeros(self._n + 1, dtype=np.int64)
        self._neg = np.zeros(self._n + 1, dtype=np.int64)

    def accumulate(self):
        tot_pos = tot_neg = 0.0
        auc = 0.0
        for i in range(self._n, -1, -1):
            np_, nn_ = tot_pos + self._pos[i], tot_neg + self._neg[i]
            auc += (nn_ - tot_neg) * (np_ + tot_pos) / 2.0
            tot_pos, tot_neg = np_, nn_
        return auc / tot_pos / tot_neg if tot_pos > 0 and tot_neg > 0 else 0.0

    def name(self):
        return self._name


def accuracy(input, label, k=1, correct=None, total=None, name=None):
    x = input.as_subclass(torch.Tensor)
    l = label.as_subclass(torch.Tensor).reshape(-1, 1)
    idx = x.topk(k, -1).indices
    return (idx == l).any(-1).float().mean().as_subclass(Tensor)
