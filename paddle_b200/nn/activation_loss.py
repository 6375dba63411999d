"""Activation and loss layers. Parity: python/paddle/nn/layer/activation.py, loss.py."""
from __future__ import annotations

from . import functional as F
from . import initializer as I
from .layer import Layer


def _act(name, fn, arg_names=(), defaults=None):
    defaults = defaults or {}

    def __init__(self, *args, **kwargs):
        Layer.__init__(self)
        cfg = dict(defaults)
        for k, v in zip(arg_names, args):
            cfg[k] = v
        for k, v in kwargs.items():
            if k != "name":
                cfg[k] = v
        self._cfg = cfg

    def forward(self, x):
        return fn(x, **self._cfg)

    def extra_repr(self):
        return ", ".join(f"{k}={v}" for k, v in self._cfg.items())

    return type(name, (Layer,), {"__init__": __init__, "forward": forward, "extra_repr": extra_repr})


ReLU = _act("ReLU", F.relu)
ReLU6 = _act("ReLU6", F.relu6)
LeakyReLU = _act("LeakyReLU", F.leaky_relu, ["negative_slope"], dict(negative_slope=0.01))
ELU = _act("ELU", F.elu, ["alpha"], dict(alpha=1.0))
CELU = _act("CELU", F.celu, ["alpha"], dict(alpha=1.0))
SELU = _act("SELU", F.selu, ["scale", "alpha"], dict(scale=1.0507009873554804934193349852946, alpha=1.6732632423543772848170429916717))
GELU = _act("GELU", F.gelu, ["approximate"], dict(approximate=False))
Silu = _act("Silu", F.silu)
Swish = _act("Swish", F.swish)
Mish = _act("Mish", F.mish)
Sigmoid = _act("Sigmoid", F.sigmoid)
Hardsigmoid = _act("Hardsigmoid", F.hardsigmoid)
Hardswish = _act("Hardswish", F.hardswish)
Hardtanh = _act("Hardtanh", F.hardtanh, ["min", "max"], dict(min=-1.0, max=1.0))
Hardshrink = _act("Hardshrink", F.hardshrink, ["threshold"], dict(threshold=0.5))
Softshrink = _act("Softshrink", F.softshrink, ["threshold"], dict(threshold=0.5))
Tanhshrink = _act("Tanhshrink", F.tanhshrink)
Tanh = _act("Tanh", F.tanh)
Softplus = _act("Softplus", F.softplus, ["beta", "threshold"], dict(beta=1, threshold=20))
Softsign = _act("Softsign", F.softsign)
LogSigmoid = _act("LogSigmoid", F.log_sigmoid)
ThresholdedReLU = _act("ThresholdedReLU", F.thresholded_relu, ["threshold", "value"], dict(threshold=1.0, value=0.0))
Maxout = _act("Maxout", F.maxout, ["groups", "axis"], dict(axis=1))
Softmax = _act("Softmax", F.softmax, ["axis"], dict(axis=-1))
LogSoftmax = _act("LogSoftmax", F.log_softmax, ["axis"], dict(axis=-1))
GLU = _act("GLU", F.glu, ["axis"], dict(axis=-1))
class RReLU(Layer):
    """Randomised leaky ReLU: slope ~ U(lower, upper) in training, (lower + upper) / 2 in eval. Parity: nn/layer/activation.py:RReLU."""

    def __init__(self, lower=1. / 8, upper=1. / 3, name=None):
        super().__init__()
        self.lower, self.upper = lower, upper

    def forward(self, x):
        return F.rrelu(x, self.lower, self.upper, self.training)


class Softmax2D(Layer):
    def forward(self, x):
        return F.softmax(x, axis=-3)


class PReLU(Layer):
    def __init__(self, num_parameters=1, init=0.25, weight_attr=None, data_format="NCHW", name=None):
        super().__init__()
        self._data_format = data_format
        self._weight = self.create_parameter([num_parameters], attr=weight_attr, default_initializer=I.Constant(init))

    def forward(self, x):
        return F.prelu(x, self._weight, self._data_format)


# ----------------------------------------------------------------------------------------------- losses
def _loss(name, fn, arg_names=(), defaults=None, n_inputs=2):
    defaults = defaults or {}

    def __init__(self, *args, **kwargs):
        Layer.__init__(self)
        cfg = dict(defaults)
        for k, v in zip(arg_names, args):
            cfg[k] = v
        for k, v in kwargs.items():
            if k != "name":
                cfg[k] = v
        self._cfg = cfg

    def forward(self, *inputs):
        return fn(*inputs, **self._cfg)

    return type(name, (Layer,), {"__init__": __init__, "forward": forward})


CrossEntropyLoss = _loss("CrossEntropyLoss", F.cross_entropy, ["weight", "ignore_index", "reduction", "soft_label", "axis", "use_softmax", "label_smoothing"],
                         dict(weight=None, ignore_index=-100, reduction="mean", soft_label=False, axis=-1, use_softmax=True, label_smoothing=0.0))
MSELoss = _loss("MSELoss", F.mse_loss, ["reduction"], dict(reduction="mean"))
L1Loss = _loss("L1Loss", F.l1_loss, ["reduction"], dict(reduction="mean"))
SmoothL1Loss = _loss("SmoothL1Loss", F.smooth_l1_loss, ["reduction", "delta"], dict(reduction="mean", delta=1.0))
HuberLoss = _loss("HuberLoss", F.huber_loss, ["delta", "reduction"], dict(delta=1.0, reduction="mean"))
BCELoss = _loss("BCELoss", F.binary_cross_entropy, ["weight", "reduction"], dict(weight=None, reduction="mean"))
BCEWithLogitsLoss = _loss("BCEWithLogitsLoss", F.binary_cross_entropy_with_logits, ["weight", "reduction", "pos_weight"], dict(weight=None, reduction="mean", pos_weight=None))
NLLLoss = _loss("NLLLoss", F.nll_loss, ["weight", "ignore_index", "reduction"], dict(weight=None, ignore_index=-100, reduction="mean"))
KLDivLoss = _loss("KLDivLoss", F.kl_div, ["reduction", "log_target"], dict(reduction="mean", log_target=False))
MarginRankingLoss = _loss("MarginRankingLoss", F.margin_ranking_loss, ["margin", "reduction"], dict(margin=0.0, reduction="mean"), 3)
HingeEmbeddingLoss = _loss("HingeEmbeddingLoss", F.hinge_embedding_loss, ["margin", "reduction"], dict(margin=1.0, reduction="mean"))
CosineEmbeddingLoss = _loss("CosineEmbeddingLoss", F.cosine_embedding_loss, ["margin", "reduction"], dict(margin=0, reduction="mean"), 3)
TripletMarginLoss = _loss("TripletMarginLoss", F.triplet_margin_loss, ["margin", "p", "epsilon", "swap", "reduction"], dict(margin=1.0, p=2.0, epsilon=1e-6, swap=False, reduction="mean"), 3)
TripletMarginWithDistanceLoss = _loss("TripletMarginWithDistanceLoss", F.triplet_margin_with_distance_loss, ["distance_function", "margin", "swap", "reduction"], dict(distance_function=None, margin=1.0, swap=False, reduction="mean"), 3)
MultiLabelSoftMarginLoss = _loss("MultiLabelSoftMarginLoss", F.multi_label_soft_margin_loss, ["weight", "reduction"], dict(weight=None, reduction="mean"))
MultiMarginLoss = _loss("MultiMarginLoss", F.multi_margin_loss, ["p", "margin", "weight", "reduction"], dict(p=1, margin=1.0, weight=None, reduction="mean"))
SoftMarginLoss = _loss("SoftMarginLoss", F.soft_margin_loss, ["reduction"], dict(reduction="mean"))
PoissonNLLLoss = _loss("PoissonNLLLoss", F.poisson_nll_loss, ["log_input", "full", "epsilon", "reduction"], dict(log_input=True, full=False, epsilon=1e-8, reduction="mean"))
GaussianNLLLoss = _loss("GaussianNLLLoss", F.gaussian_nll_loss, ["full", "epsilon", "reduction"], dict(full=False, epsilon=1e-6, reduction="mean"), 3)
CTCLoss = _loss("CTCLoss", lambda lp, lab, il, ll, norm_by_times=False, blank=0, reduction="mean": F.ctc_loss(lp, lab, il, ll, blank, reduction, norm_by_times), ["blank", "reduction"], dict(blank=0, reduction="mean"), 4)
RNNTLoss = _loss("RNNTLoss", F.rnnt_loss, ["blank", "fastemit_lambda", "reduction"], dict(blank=0, fastemit_lambda=0.001, reduction="mean"), 4)


class HSigmoidLoss(Layer):
    def __init__(self, feature_size, num_classes, weight_attr=None, bias_attr=None, is_custom=False, is_sparse=False, name=None):
        super().__init__()
        self._num_classes, self._is_custom = num_classes, is_custom
        rows = num_classes if is_custom else num_classes - 1
        self.weight = self.create_parameter([rows, feature_size], attr=weight_attr)
        self.bias = self.create_parameter([rows, 1], attr=bias_attr, is_bias=True)

    def forward(self, input, label, path_table=None, path_code=None):
        return F.hsigmoid_loss(input, label, self._num_classes, self.weight, self.bias, path_table, path_code)


class AdaptiveLogSoftmaxWithLoss(Layer):
    def __init__(self, in_features, n_classes, cutoffs, weight_attr=None, bias_attr=None, div_value=4.0, head_bias=False, name=None):
        super().__init__()
        self.cutoffs = list(cutoffs) + [n_classes]
        self.shortlist = self.cutoffs[0]
        self.n_clusters = len(self.cutoffs) - 1
        self.head_weight = self.create_parameter([in_features, self.shortlist + self.n_clusters], attr=weight_attr)
        self.head_bias = self.create_parameter([self.shortlist + self.n_clusters], attr=bias_attr, is_bias=True) if head_bias else None
        self.tail_weights = []
        for i in range(self.n_clusters):
            hsz = max(1, int(in_features // (div_value ** (i + 1))))
            osz = self.cutoffs[i + 1] - self.cutoffs[i]
            p = self.create_parameter([in_features, hsz], attr=weight_attr)
            c = self.create_parameter([hsz, osz], attr=weight_attr)
            self.add_parameter(f"tail_proj_{i}", p)
            self.add_parameter(f"tail_cls_{i}", c)
            self.tail_weights.append((p, c))

    def forward(self, input, label):
        return F.adaptive_log_softmax_with_loss(input, label, self.head_weight, self.tail_weights, self.cutoffs, self.head_bias)

    def log_prob(self, input):
        """Full [N, n_classes] log-probabilities."""
        import torch

        x = input
        head = x @ self.head_weight + (self.head_bias if self.head_bias is not None else 0)
        head_lp = F.log_softmax(head, -1)
        parts = [head_lp[:, :self.shortlist]]
        for i, (p, c) in enumerate(self.tail_weights):
            tail_lp = F.log_softmax((x @ p) @ c, -1)
            parts.append(tail_lp + head_lp[:, self.shortlist + i:self.shortlist + i + 1])
        return torch.cat(parts, -1)

    def predict(self, input):
        return self.log_prob(input).argmax(-1)
