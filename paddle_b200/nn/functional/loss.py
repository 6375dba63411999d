"""Loss functionals. Parity: python/paddle/nn/functional/loss.py."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ...ops._helpers import T, raw, wrap


def _reduce(loss, reduction):
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def cross_entropy(input, label, weight=None, ignore_index=-100, reduction="mean", soft_label=False, axis=-1,
                  use_softmax=True, label_smoothing=0.0, name=None):
    """Parity: nn/functional/loss.py:cross_entropy (phi cross_entropy_with_softmax kernel)."""
    x, label = T(input), T(label)
    axis = axis % x.dim()
    if axis != x.dim() - 1:
        x = torch.movedim(x, axis, -1)
        if soft_label or label.dim() == x.dim():
            label = torch.movedim(label, axis, -1)
    if soft_label or (label.dim() == x.dim() and label.size(-1) == x.size(-1) and label.is_floating_point()):
        logp = F.log_softmax(x.float(), -1) if use_softmax else torch.log(x.float())
        lab = label.float()
        if label_smoothing > 0:
            lab = lab * (1 - label_smoothing) + label_smoothing / x.size(-1)
        loss = -(lab * logp).sum(-1)
        if weight is not None:
            loss = loss * (lab * T(weight).float()).sum(-1)
        return _reduce(loss, reduction).to(x.dtype) if reduction != "none" else loss.to(x.dtype)
    if label.dim() == x.dim():
        label = label.squeeze(-1)
    label = label.long()
    if use_softmax and x.is_cuda and weight is None and label_smoothing == 0.0:
        from ...kernels import loss as K

        per = K.softmax_cross_entropy(x.reshape(-1, x.size(-1)), label.reshape(-1), ignore_index).reshape_as(label)     # reshape_as: no extents baked into a recorded program
        if reduction == "none":
            return per
        if reduction == "sum":
            return per.sum()
        valid = (label != ignore_index).sum().clamp(min=1)
        return per.sum() / valid.to(per.dtype)
    lead = x.reshape(-1, x.size(-1))
    lab = label.reshape(-1)
    if use_softmax:
        loss = F.cross_entropy(lead.float(), lab, None if weight is None else T(weight).float(), ignore_index=ignore_index,
                               reduction="none", label_smoothing=label_smoothing)
    else:
        loss = F.nll_loss(torch.log(lead.float()), lab, None if weight is None else T(weight).float(), ignore_index=ignore_index, reduction="none")
    loss = loss.reshape_as(label)          # by reference to the label: a recorded program keeps working when the batch extent is dynamic
    if reduction == "none":
        return loss.to(x.dtype)
    if reduction == "sum":
        return loss.sum().to(x.dtype)
    mask = lab != ignore_index
    if weight is not None:
        denom = (T(weight).float()[lab.clamp(min=0)] * mask).sum()
    else:
        denom = mask.sum()
    return (loss.sum() / denom.clamp(min=1e-12)).to(x.dtype)


def softmax_with_cross_entropy(logits, label, soft_label=False, ignore_index=-100, numeric_stable_mode=True,
                               return_softmax=False, axis=-1):
    loss = cross_entropy(logits, label, ignore_index=ignore_index, reduction="none", soft_label=soft_label, axis=axis)
    loss = loss.unsqueeze(axis)
    if return_softmax:
        return loss, F.softmax(T(logits), axis)
    return loss


def nll_loss(input, label, weight=None, ignore_index=-100, reduction="mean", name=None):
    return F.nll_loss(T(input), T(label).long(), None if weight is None else T(weight), ignore_index=ignore_index, reduction=reduction)


def mse_loss(input, label, reduction="mean", name=None):
    return F.mse_loss(T(input), T(label), reduction=reduction)


def l1_loss(input, label, reduction="mean", name=None):
    return F.l1_loss(T(input), T(label), reduction=reduction)


def smooth_l1_loss(input, label, reduction="mean", delta=1.0, name=None):
    # paddle's smooth_l1 is huber with delta
    return F.huber_loss(T(input), T(label), reduction=reduction, delta=delta)


def huber_loss(input, label, delta=1.0, reduction="mean", name=None):
    return F.huber_loss(T(input), T(label), reduction=reduction, delta=delta)


def binary_cross_entropy(input, label, weight=None, reduction="mean", name=None):
    return F.binary_cross_entropy(T(input), T(label), None if weight is None else T(weight), reduction=reduction)


def binary_cross_entropy_with_logits(logit, label, weight=None, reduction="mean", pos_weight=None, name=None):
    return F.binary_cross_entropy_with_logits(T(logit), T(label), None if weight is None else T(weight), reduction=reduction,
                                              pos_weight=None if pos_weight is None else T(pos_weight))


def kl_div(input, label, reduction="mean", log_target=False, name=None):
    return F.kl_div(T(input), T(label), reduction=reduction, log_target=log_target)


def margin_ranking_loss(input, other, label, margin=0.0, reduction="mean", name=None):
    return F.margin_ranking_loss(T(input), T(other), T(label), margin=margin, reduction=reduction)


def hinge_embedding_loss(input, label, margin=1.0, reduction="mean", name=None):
    return F.hinge_embedding_loss(T(input), T(label), margin=margin, reduction=reduction)


def cosine_embedding_loss(input1, input2, label, margin=0, reduction="mean", name=None):
    return F.cosine_embedding_loss(T(input1), T(input2), T(label), margin=margin, reduction=reduction)


def triplet_margin_loss(input, positive, negative, margin=1.0, p=2, epsilon=1e-06, swap=False, reduction="mean", name=None):
    return F.triplet_margin_loss(T(input), T(positive), T(negative), margin=margin, p=p, eps=epsilon, swap=swap, reduction=reduction)


def triplet_margin_with_distance_loss(input, positive, negative, distance_function=None, margin=1.0, swap=False, reduction="mean", name=None):
    return F.triplet_margin_with_distance_loss(T(input), T(positive), T(negative), distance_function=distance_function, margin=margin, swap=swap, reduction=reduction)


def multi_label_soft_margin_loss(input, label, weight=None, reduction="mean", name=None):
    return F.multilabel_soft_margin_loss(T(input), T(label), None if weight is None else T(weight), reduction=reduction)


def multi_margin_loss(input, label, p=1, margin=1.0, weight=None, reduction="mean", name=None):
    return F.multi_margin_loss(T(input), T(label).long(), p=p, margin=margin, weight=None if weight is None else T(weight), reduction=reduction)


def soft_margin_loss(input, label, reduction="mean", name=None):
    return F.soft_margin_loss(T(input), T(label).to(input.dtype), reduction=reduction)


def poisson_nll_loss(input, label, log_input=True, full=False, epsilon=1e-8, reduction="mean", name=None):
    return F.poisson_nll_loss(T(input), T(label), log_input=log_input, full=full, eps=epsilon, reduction=reduction)


def gaussian_nll_loss(input, label, variance, full=False, epsilon=1e-6, reduction="mean", name=None):
    return F.gaussian_nll_loss(T(input), T(label), T(variance), full=full, eps=epsilon, reduction=reduction)


def log_loss(input, label, epsilon=1e-4, name=None):
    x, y = T(input), T(label)
    return -y * torch.log(x + epsilon) - (1 - y) * torch.log(1 - x + epsilon)


def square_error_cost(input, label):
    return (T(input) - T(label)) ** 2


def sigmoid_focal_loss(logit, label, normalizer=None, alpha=0.25, gamma=2.0, reduction="sum", name=None):
    logit, label = T(logit), T(label)
    p = torch.sigmoid(logit)
    ce = F.binary_cross_entropy_with_logits(logit, label, reduction="none")
    pt = p * label + (1 - p) * (1 - label)
    loss = ce * (1 - pt) ** gamma
    if alpha >= 0:
        loss = (alpha * label + (1 - alpha) * (1 - label)) * loss
    if normalizer is not None:
        loss = loss / T(normalizer)
    return _reduce(loss, reduction)


def dice_loss(input, label, epsilon=1e-5, name=None):
    x, label = T(input), T(label)
    oh = F.one_hot(label.squeeze(-1).long(), x.size(-1)).to(x.dtype)
    dims = tuple(range(1, x.dim()))
    inter = (x * oh).sum(dims)
    return (1 - 2 * inter / (x.sum(dims) + oh.sum(dims) + epsilon)).mean()


def npair_loss(anchor, positive, labels, l2_reg=0.002):
    a, p, labels = T(anchor), T(positive), T(labels).reshape(-1, 1).float()
    eq = (labels == labels.t()).float()
    tgt = eq / eq.sum(1, keepdim=True)
    l2 = ((a ** 2).sum(1).mean() + (p ** 2).sum(1).mean()) * 0.25 * l2_reg
    sim = a @ p.t()
    ce = (-tgt * F.log_softmax(sim, 1)).sum(1).mean()
    return l2 + ce


def ctc_loss(log_probs, labels, input_lengths, label_lengths, blank=0, reduction="mean", norm_by_times=False):
    lp = F.log_softmax(T(log_probs).float(), -1)
    loss = F.ctc_loss(lp, T(labels).long(), T(input_lengths).long(), T(label_lengths).long(), blank=blank, reduction="none", zero_infinity=False)
    if reduction == "mean":
        return (loss / T(label_lengths).to(loss.dtype)).mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def rnnt_loss(input, label, input_lengths, label_lengths, blank=0, fastemit_lambda=0.001, reduction="mean", name=None):
    """RNN-T loss by forward-variable DP (log-space). Parity: nn/functional/loss.py:rnnt_loss (warprnnt)."""
    x = F.log_softmax(T(input).float(), -1)  # [B, T, U+1, V]
    B, Tm, U1, V = x.shape
    losses = []
    for b in range(B):
        t_len, u_len = int(input_lengths[b]), int(label_lengths[b])
        lab = T(label)[b, :u_len].long()
        blank_lp = x[b, :t_len, : u_len + 1, blank]
        emit_lp = torch.gather(x[b, :t_len, :u_len, :], 2, lab.reshape(1, -1, 1).expand(t_len, u_len, 1)).squeeze(-1) if u_len > 0 else None
        alpha = [[None] * (u_len + 1) for _ in range(t_len)]
        alpha[0][0] = x.new_zeros(())
        for t in range(t_len):
            for u in range(u_len + 1):
                if t == 0 and u == 0:
                    continue
                terms = []
                if t > 0:
                    terms.append(alpha[t - 1][u] + blank_lp[t - 1, u])
                if u > 0:
                    terms.append(alpha[t][u - 1] + emit_lp[t, u - 1])
                alpha[t][u] = torch.logsumexp(torch.stack(terms), 0)
        losses.append(-(alpha[t_len - 1][u_len] + blank_lp[t_len - 1, u_len]))
    loss = torch.stack(losses)
    return _reduce(loss, reduction)


def margin_cross_entropy(logits, label, margin1=1.0, margin2=0.5, margin3=0.0, scale=64.0, group=None,
                         return_softmax=False, reduction="mean"):
    x, label = T(logits).float(), T(label).long().reshape(-1)
    theta = torch.acos(x.clamp(-1 + 1e-7, 1 - 1e-7))
    tgt = torch.cos(margin1 * theta + margin2) - margin3
    oh = F.one_hot(label, x.size(-1)).bool()
    out = torch.where(oh, tgt, x) * scale
    loss = F.cross_entropy(out, label, reduction="none").unsqueeze(-1)
    loss = _reduce(loss, reduction)
    return (loss, F.softmax(out, -1)) if return_softmax else loss


def hsigmoid_loss(input, label, num_classes, weight, bias=None, path_table=None, path_code=None, is_sparse=False, name=None):
    """Default complete-binary-tree hierarchical sigmoid. Parity: nn/functional/loss.py:hsigmoid_loss."""
    x, label, w = T(input), T(label).long().reshape(-1), T(weight)
    if path_table is not None:
        table, code = T(path_table).long(), T(path_code).float()
        valid = (table >= 0).float()
        wsel = w[table.clamp(min=0)]
        logits = torch.einsum("bd,bld->bl", x, wsel)
        if bias is not None:
            logits = logits + T(bias).reshape(-1)[table.clamp(min=0)]
        loss = F.binary_cross_entropy_with_logits(logits, code, reduction="none") * valid
        return loss.sum(1, keepdim=True)
    # default tree: class c is leaf (c + num_classes) of the complete binary tree; its path is c >> 1, c >> 2, ... down to the root, the code bits
    # are the low bits on the way - built for the whole batch at once
    depth = max(1, (2 * num_classes - 1).bit_length())
    c = (label.reshape(-1).long() + num_classes)[:, None]
    d = torch.arange(depth, device=x.device)[None]
    cur = c >> d                                           # node code at every level, leaf first
    valid = (cur > 1).to(x.dtype)
    table = ((cur >> 1) - 1).clamp(min=0)
    code = (cur & 1).to(x.dtype)
    logits = torch.einsum("bd,bld->bl", x, w[table])
    if bias is not None:
        logits = logits + T(bias).reshape(-1)[table]
    loss = F.binary_cross_entropy_with_logits(logits, code, reduction="none") * valid
    return loss.sum(1, keepdim=True)


def adaptive_log_softmax_with_loss(input, label, head_weight, tail_weights, cutoffs, head_bias=None, name=None):
    x, label = T(input), T(label).long()
    n_clusters = len(cutoffs) - 1 if cutoffs[-1] >= 0 else len(cutoffs)
    cut = list(cutoffs)
    shortlist = cut[0]
    head = x @ T(head_weight) + (T(head_bias) if head_bias is not None else 0)
    head_lp = F.log_softmax(head, -1)
    out = x.new_zeros(x.size(0))
    in_short = label < shortlist
    out = torch.where(in_short, torch.gather(head_lp, 1, label.clamp(max=shortlist - 1).unsqueeze(1)).squeeze(1), out)
    for i, (lo, hi) in enumerate(zip(cut[:-1], cut[1:])):
        mask = (label >= lo) & (label < hi)
        if not bool(mask.any()):
            continue
        proj, cls = tail_weights[i]
        tail_lp = F.log_softmax((x @ T(proj)) @ T(cls), -1)
        rel = (label - lo).clamp(0, hi - lo - 1)
        lp = head_lp[:, shortlist + i] + torch.gather(tail_lp, 1, rel.unsqueeze(1)).squeeze(1)
        out = torch.where(mask, lp, out)
    return out, -out.mean()


__all__ = [n for n in list(globals()) if not n.startswith("_") and n not in ("torch", "F", "T", "raw", "wrap", "annotations")]
