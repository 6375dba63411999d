"""LeNet, AlexNet, VGG, SqueezeNet, MobileNetV1/V2/V3, ShuffleNetV2, DenseNet, GoogLeNet, InceptionV3.
Parity: python/paddle/vision/models/{lenet,alexnet,vgg,squeezenet,mobilenetv1,mobilenetv2,mobilenetv3,shufflenetv2,densenet,
googlenet,inceptionv3}.py (architectures; no pretrained download)."""
import math

import torch

from ... import nn
from ...nn import functional as F


def _no_pretrained(pretrained):
    if pretrained:
        raise RuntimeError("pretrained weights need network access; load a local .pdparams instead")


class LeNet(nn.Layer):
    def __init__(self, num_classes=10):
        super().__init__()
        self.num_classes = num_classes
        self.features = nn.Sequential(nn.Conv2D(1, 6, 3, stride=1, padding=1), nn.ReLU(), nn.MaxPool2D(2, 2),
                                      nn.Conv2D(6, 16, 5, stride=1, padding=0), nn.ReLU(), nn.MaxPool2D(2, 2))
        if num_classes > 0:
            self.fc = nn.Sequential(nn.Linear(400, 120), nn.Linear(120, 84), nn.Linear(84, num_classes))

    def forward(self, x):
        x = self.features(x)
        if self.num_classes > 0:
            x = self.fc(x.flatten(1))
        return x


class AlexNet(nn.Layer):
    def __init__(self, num_classes=1000):
        super().__init__()
        self.num_classes = num_classes
        self.features = nn.Sequential(
            nn.Conv2D(3, 64, 11, stride=4, padding=2), nn.ReLU(), nn.MaxPool2D(3, 2),
            nn.Conv2D(64, 192, 5, padding=2), nn.ReLU(), nn.MaxPool2D(3, 2),
            nn.Conv2D(192, 384, 3, padding=1), nn.ReLU(), nn.Conv2D(384, 256, 3, padding=1), nn.ReLU(),
            nn.Conv2D(256, 256, 3, padding=1), nn.ReLU(), nn.MaxPool2D(3, 2))
        self.avgpool = nn.AdaptiveAvgPool2D((6, 6))
        if num_classes > 0:
            self.classifier = nn.Sequential(nn.Dropout(0.5), nn.Linear(256 * 36, 4096), nn.ReLU(), nn.Dropout(0.5), nn.Linear(4096, 4096),
                                            nn.ReLU(), nn.Linear(4096, num_classes))

    def forward(self, x):
        x = self.avgpool(self.features(x))
        if self.num_classes > 0:
            x = self.classifier(x.flatten(1))
        return x


def alexnet(pretrained=False, **kw):
    _no_pretrained(pretrained)
    return AlexNet(**kw)


_VGG_CFG = {"A": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
            "B": [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
            "D": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
            "E": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]}


class VGG(nn.Layer):
    def __init__(self, features, num_classes=1000, with_pool=True):
        super().__init__()
        self.features, self.num_classes, self.with_pool = features, num_classes, with_pool
        if with_pool:
            self.avgpool = nn.AdaptiveAvgPool2D((7, 7))
        if num_classes > 0:
            self.classifier = nn.Sequential(nn.Linear(512 * 49, 4096), nn.ReLU(), nn.Dropout(), nn.Linear(4096, 4096), nn.ReLU(), nn.Dropout(),
                                            nn.Linear(4096, num_classes))

    def forward(self, x):
        x = self.features(x)
        if self.with_pool:
            x = self.avgpool(x)
        if self.num_classes > 0:
            x = self.classifier(x.flatten(1))
        return x


def _vgg_layers(cfg, batch_norm):
    layers, c = [], 3
    for v in cfg:
        if v == "M":
            layers.append(nn.MaxPool2D(2, 2))
        else:
            layers.append(nn.Conv2D(c, v, 3, padding=1))
            if batch_norm:
                layers.append(nn.BatchNorm2D(v))
            layers.append(nn.ReLU())
            c = v
    return nn.Sequential(*layers)


def vgg11(pretrained=False, batch_norm=False, **kw):
    _no_pretrained(pretrained)
    return VGG(_vgg_layers(_VGG_CFG["A"], batch_norm), **kw)


def vgg13(pretrained=False, batch_norm=False, **kw):
    _no_pretrained(pretrained)
    return VGG(_vgg_layers(_VGG_CFG["B"], batch_norm), **kw)


def vgg16(pretrained=False, batch_norm=False, **kw):
    _no_pretrained(pretrained)
    return VGG(_vgg_layers(_VGG_CFG["D"], batch_norm), **kw)


def vgg19(pretrained=False, batch_norm=False, **kw):
    _no_pretrained(pretrained)
    return VGG(_vgg_layers(_VGG_CFG["E"], batch_norm), **kw)


class _Fire(nn.Layer):
    def __init__(self, inp, sq, e1, e3):
        super().__init__()
        self.squeeze = nn.Conv2D(inp, sq, 1)
        self.e1 = nn.Conv2D(sq, e1, 1)
        self.e3 = nn.Conv2D(sq, e3, 3, padding=1)

    def forward(self, x):
        x = F.relu(self.squeeze(x))
        return torch.cat([F.relu(self.e1(x)), F.relu(self.e3(x))], 1)


class SqueezeNet(nn.Layer):
    def __init__(self, version="1.0", num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        if version == "1.0":
            self.features = nn.Sequential(nn.Conv2D(3, 96, 7, stride=2), nn.ReLU(), nn.MaxPool2D(3, 2), _Fire(96, 16, 64, 64), _Fire(128, 16, 64, 64),
                                          _Fire(128, 32, 128, 128), nn.MaxPool2D(3, 2), _Fire(256, 32, 128, 128), _Fire(256, 48, 192, 192),
                                          _Fire(384, 48, 192, 192), _Fire(384, 64, 256, 256), nn.MaxPool2D(3, 2), _Fire(512, 64, 256, 256))
        else:
            self.features = nn.Sequential(nn.Conv2D(3, 64, 3, stride=2, padding=1), nn.ReLU(), nn.MaxPool2D(3, 2), _Fire(64, 16, 64, 64),
                                          _Fire(128, 16, 64, 64), nn.MaxPool2D(3, 2), _Fire(128, 32, 128, 128), _Fire(256, 32, 128, 128),
                                          nn.MaxPool2D(3, 2), _Fire(256, 48, 192, 192), _Fire(384, 48, 192, 192), _Fire(384, 64, 256, 256),
                                          _Fire(512, 64, 256, 256))
        if num_classes > 0:
            self.classifier = nn.Sequential(nn.Dropout(0.5), nn.Conv2D(512, num_classes, 1), nn.ReLU())
        if with_pool:
            self.avgpool = nn.AdaptiveAvgPool2D((1, 1))

    def forward(self, x):
        x = self.features(x)
        if self.num_classes > 0:
            x = self.classifier(x)
        if self.with_pool:
            x = self.avgpool(x)
        return x.flatten(1) if self.num_classes > 0 else x


def squeezenet1_0(pretrained=False, **kw):
    _no_pretrained(pretrained)
    return SqueezeNet("1.0", **kw)


def squeezenet1_1(pretrained=False, **kw):
    _no_pretrained(pretrained)
    return SqueezeNet("1.1", **kw)


def _cbr(i, o, k, s=1, p=0, g=1, act="relu"):
    layers = [nn.Conv2D(i, o, k, stride=s, padding=p, groups=g, bias_attr=False), nn.BatchNorm2D(o)]
    if act == "relu":
        layers.append(nn.ReLU())
    elif act == "relu6":
        layers.append(nn.ReLU6())
    elif act == "hardswish":
        layers.append(nn.Hardswish())
    return nn.Sequential(*layers)


class MobileNetV1(nn.Layer):
    def __init__(self, scale=1.0, num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        c = lambda v: int(v * scale)  # noqa: E731
        cfg = [(32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2)] + [(512, 512, 1)] * 5 + [(512, 1024, 2), (1024, 1024, 1)]
        layers = [_cbr(3, c(32), 3, 2, 1)]
        for i, o, s in cfg:
            layers += [_cbr(c(i), c(i), 3, s, 1, g=c(i)), _cbr(c(i), c(o), 1)]
        self.features = nn.Sequential(*layers)
        if with_pool:
            self.pool2d_avg = nn.AdaptiveAvgPool2D(1)
        if num_classes > 0:
            self.fc = nn.Linear(c(1024), num_classes)

    def forward(self, x):
        x = self.features(x)
        if self.with_pool:
            x = self.pool2d_avg(x)
        if self.num_classes > 0:
            x = self.fc(x.flatten(1))
        return x


def mobilenet_v1(pretrained=False, scale=1.0, **kw):
    _no_pretrained(pretrained)
    return MobileNetV1(scale=scale, **kw)


def _make_divisible(v, divisor=8, min_value=None):
    min_value = min_value or divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    return new_v + divisor if new_v < 0.9 * v else new_v


class _InvertedResidual(nn.Layer):
    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        hidden = int(round(inp * expand_ratio))
        self.use_res = stride == 1 and inp == oup
        layers = []
        if expand_ratio != 1:
            layers.append(_cbr(inp, hidden, 1, act="relu6"))
        layers += [_cbr(hidden, hidden, 3, stride, 1, g=hidden, act="relu6"), nn.Conv2D(hidden, oup, 1, bias_attr=False), nn.BatchNorm2D(oup)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.conv(x) if self.use_res else self.conv(x)


class MobileNetV2(nn.Layer):
    def __init__(self, scale=1.0, num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        inp = _make_divisible(32 * scale)
        last = _make_divisible(1280 * max(1.0, scale))
        cfg = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
        feats = [_cbr(3, inp, 3, 2, 1, act="relu6")]
        for t, c, n, s in cfg:
            oup = _make_divisible(c * scale)
            for i in range(n):
                feats.append(_InvertedResidual(inp, oup, s if i == 0 else 1, t))
                inp = oup
        feats.append(_cbr(inp, last, 1, act="relu6"))
        self.features = nn.Sequential(*feats)
        if with_pool:
            self.pool2d_avg = nn.AdaptiveAvgPool2D(1)
        if num_classes > 0:
            self.classifier = nn.Sequential(nn.Dropout(0.2), nn.Linear(last, num_classes))

    def forward(self, x):
        x = self.features(x)
        if self.with_pool:
            x = self.pool2d_avg(x)
        if self.num_classes > 0:
            x = self.classifier(x.flatten(1))
        return x


def mobilenet_v2(pretrained=False, scale=1.0, **kw):
    _no_pretrained(pretrained)
    return MobileNetV2(scale=scale, **kw)


class _SE(nn.Layer):
    def __init__(self, c, r=4):
        super().__init__()
        self.fc1 = nn.Conv2D(c, _make_divisible(c // r), 1)
        self.fc2 = nn.Conv2D(_make_divisible(c // r), c, 1)

    def forward(self, x):
        s = F.adaptive_avg_pool2d(x, 1)
        return x * F.hardsigmoid(self.fc2(F.relu(self.fc1(s))), slope=0.2, offset=0.5)


class _MBV3Block(nn.Layer):
    def __init__(self, inp, k, exp, out, se, act, s):
        super().__init__()
        self.use_res = s == 1 and inp == out
        layers = []
        if exp != inp:
            layers.append(_cbr(inp, exp, 1, act=act))
        layers.append(_cbr(exp, exp, k, s, (k - 1) // 2, g=exp, act=act))
        if se:
            layers.append(_SE(exp))
        layers += [nn.Conv2D(exp, out, 1, bias_attr=False), nn.BatchNorm2D(out)]
        self.block = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.block(x) if self.use_res else self.block(x)


class MobileNetV3(nn.Layer):
    def __init__(self, config, last_channel, scale=1.0, num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        inp = _make_divisible(16 * scale)
        feats = [_cbr(3, inp, 3, 2, 1, act="hardswish")]
        for k, exp, out, se, act, s in config:
            e, o = _make_divisible(exp * scale), _make_divisible(out * scale)
            feats.append(_MBV3Block(inp, k, e, o, se, act, s))
            inp = o
        lastconv = _make_divisible(6 * inp)
        feats.append(_cbr(inp, lastconv, 1, act="hardswish"))
        self.features = nn.Sequential(*feats)
        if with_pool:
            self.avgpool = nn.AdaptiveAvgPool2D(1)
        if num_classes > 0:
            self.classifier = nn.Sequential(nn.Linear(lastconv, last_channel), nn.Hardswish(), nn.Dropout(0.2), nn.Linear(last_channel, num_classes))

    def forward(self, x):
        x = self.features(x)
        if self.with_pool:
            x = self.avgpool(x)
        if self.num_classes > 0:
            x = self.classifier(x.flatten(1))
        return x


_MBV3_SMALL = [(3, 16, 16, True, "relu", 2), (3, 72, 24, False, "relu", 2), (3, 88, 24, False, "relu", 1), (5, 96, 40, True, "hardswish", 2),
               (5, 240, 40, True, "hardswish", 1), (5, 240, 40, True, "hardswish", 1), (5, 120, 48, True, "hardswish", 1),
               (5, 144, 48, True, "hardswish", 1), (5, 288, 96, True, "hardswish", 2), (5, 576, 96, True, "hardswish", 1), (5, 576, 96, True, "hardswish", 1)]
_MBV3_LARGE = [(3, 16, 16, False, "relu", 1), (3, 64, 24, False, "relu", 2), (3, 72, 24, False, "relu", 1), (5, 72, 40, True, "relu", 2),
               (5, 120, 40, True, "relu", 1), (5, 120, 40, True, "relu", 1), (3, 240, 80, False, "hardswish", 2), (3, 200, 80, False, "hardswish", 1),
               (3, 184, 80, False, "hardswish", 1), (3, 184, 80, False, "hardswish", 1), (3, 480, 112, True, "hardswish", 1),
               (3, 672, 112, True, "hardswish", 1), (5, 672, 160, True, "hardswish", 2), (5, 960, 160, True, "hardswish", 1), (5, 960, 160, True, "hardswish", 1)]


class MobileNetV3Small(MobileNetV3):
    def __init__(self, scale=1.0, num_classes=1000, with_pool=True):
        super().__init__(_MBV3_SMALL, _make_divisible(1024 * scale), scale, num_classes, with_pool)


class MobileNetV3Large(MobileNetV3):
    def __init__(self, scale=1.0, num_classes=1000, with_pool=True):
        super().__init__(_MBV3_LARGE, _make_divisible(1280 * scale), scale, num_classes, with_pool)


def mobilenet_v3_small(pretrained=False, scale=1.0, **kw):
    _no_pretrained(pretrained)
    return MobileNetV3Small(scale=scale, **kw)


def mobilenet_v3_large(pretrained=False, scale=1.0, **kw):
    _no_pretrained(pretrained)
    return MobileNetV3Large(scale=scale, **kw)


class _ShuffleUnit(nn.Layer):
    def __init__(self, inp, oup, stride, act="relu"):
        super().__init__()
        self.stride = stride
        bf = oup // 2
        if stride > 1:
            self.branch1 = nn.Sequential(nn.Conv2D(inp, inp, 3, stride=stride, padding=1, groups=inp, bias_attr=False), nn.BatchNorm2D(inp),
                                         _cbr(inp, bf, 1, act=act))
        self.branch2 = nn.Sequential(_cbr(inp if stride > 1 else bf, bf, 1, act=act),
                                     nn.Conv2D(bf, bf, 3, stride=stride, padding=1, groups=bf, bias_attr=False), nn.BatchNorm2D(bf), _cbr(bf, bf, 1, act=act))

    def forward(self, x):
        if self.stride == 1:
            x1, x2 = torch.chunk(x, 2, 1)
            out = torch.cat([x1, self.branch2(x2)], 1)
        else:
            out = torch.cat([self.branch1(x), self.branch2(x)], 1)
        return F.channel_shuffle(out, 2)


class ShuffleNetV2(nn.Layer):
    def __init__(self, scale=1.0, act="relu", num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        chans = {0.25: [24, 24, 48, 96, 512], 0.33: [24, 32, 64, 128, 512], 0.5: [24, 48, 96, 192, 1024], 1.0: [24, 116, 232, 464, 1024],
                 1.5: [24, 176, 352, 704, 1024], 2.0: [24, 244, 488, 976, 2048]}[scale]
        self.conv1 = _cbr(3, chans[0], 3, 2, 1, act=act)
        self.maxpool = nn.MaxPool2D(3, 2, 1)
        stages, inp = [], chans[0]
        for reps, oup in zip([4, 8, 4], chans[1:4]):
            stages.append(_ShuffleUnit(inp, oup, 2, act))
            stages += [_ShuffleUnit(oup, oup, 1, act) for _ in range(reps - 1)]
            inp = oup
        self.stages = nn.Sequential(*stages)
        self.conv5 = _cbr(inp, chans[4], 1, act=act)
        if with_pool:
            self.pool = nn.AdaptiveAvgPool2D(1)
        if num_classes > 0:
            self.fc = nn.Linear(chans[4], num_classes)

    def forward(self, x):
        x = self.conv5(self.stages(self.maxpool(self.conv1(x))))
        if self.with_pool:
            x = self.pool(x)
        if self.num_classes > 0:
            x = self.fc(x.flatten(1))
        return x


def _shuffle(scale, act="relu"):
    def f(pretrained=False, **kw):
        _no_pretrained(pretrained)
        return ShuffleNetV2(scale=scale, act=act, **kw)

    return f


shufflenet_v2_x0_25, shufflenet_v2_x0_33, shufflenet_v2_x0_5 = _shuffle(0.25), _shuffle(0.33), _shuffle(0.5)
shufflenet_v2_x1_0, shufflenet_v2_x1_5, shufflenet_v2_x2_0 = _shuffle(1.0), _shuffle(1.5), _shuffle(2.0)
shufflenet_v2_swish = _shuffle(1.0, "hardswish")


class _DenseLayer(nn.Layer):
    def __init__(self, inp, growth, bn_size, dropout):
        super().__init__()
        self.fn = nn.Sequential(nn.BatchNorm2D(inp), nn.ReLU(), nn.Conv2D(inp, bn_size * growth, 1, bias_attr=False),
                                nn.BatchNorm2D(bn_size * growth), nn.ReLU(), nn.Conv2D(bn_size * growth, growth, 3, padding=1, bias_attr=False))
        self.dropout = dropout

    def forward(self, x):
        y = self.fn(x)
        if self.dropout:
            y = F.dropout(y, self.dropout, training=self.training)
        return torch.cat([x, y], 1)


class DenseNet(nn.Layer):
    def __init__(self, layers=121, bn_size=4, dropout=0.0, num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        spec = {121: (64, 32, [6, 12, 24, 16]), 161: (96, 48, [6, 12, 36, 24]), 169: (64, 32, [6, 12, 32, 32]), 201: (64, 32, [6, 12, 48, 32]),
                264: (64, 32, [6, 12, 64, 48])}[layers]
        c, growth, blocks = spec
        feats = [nn.Conv2D(3, c, 7, stride=2, padding=3, bias_attr=False), nn.BatchNorm2D(c), nn.ReLU(), nn.MaxPool2D(3, 2, 1)]
        for i, n in enumerate(blocks):
            for _ in range(n):
                feats.append(_DenseLayer(c, growth, bn_size, dropout))
                c += growth
            if i != len(blocks) - 1:
                feats += [nn.BatchNorm2D(c), nn.ReLU(), nn.Conv2D(c, c // 2, 1, bias_attr=False), nn.AvgPool2D(2, 2)]
                c //= 2
        feats += [nn.BatchNorm2D(c), nn.ReLU()]
        self.features = nn.Sequential(*feats)
        if with_pool:
            self.pool = nn.AdaptiveAvgPool2D(1)
        if num_classes > 0:
            self.out = nn.Linear(c, num_classes)

    def forward(self, x):
        x = self.features(x)
        if self.with_pool:
            x = self.pool(x)
        if self.num_classes > 0:
            x = self.out(x.flatten(1))
        return x


def _dense(n):
    def f(pretrained=False, **kw):
        _no_pretrained(pretrained)
        return DenseNet(layers=n, **kw)

    return f


densenet121, densenet161, densenet169, densenet201, densenet264 = _dense(121), _dense(161), _dense(169), _dense(201), _dense(264)


class _Inception(nn.Layer):
    def __init__(self, inp, c1, c3r, c3, c5r, c5, proj):
        super().__init__()
        self.b1 = nn.Conv2D(inp, c1, 1)
        self.b2 = nn.Sequential(nn.Conv2D(inp, c3r, 1), nn.ReLU(), nn.Conv2D(c3r, c3, 3, padding=1))
        self.b3 = nn.Sequential(nn.Conv2D(inp, c5r, 1), nn.ReLU(), nn.Conv2D(c5r, c5, 5, padding=2))
        self.b4 = nn.Sequential(nn.MaxPool2D(3, 1, 1), nn.Conv2D(inp, proj, 1))

    def forward(self, x):
        return F.relu(torch.cat([self.b1(x), self.b2(x), self.b3(x), self.b4(x)], 1))


class GoogLeNet(nn.Layer):
    def __init__(self, num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        self.stem = nn.Sequential(nn.Conv2D(3, 64, 7, stride=2, padding=3), nn.ReLU(), nn.MaxPool2D(3, 2, 1), nn.Conv2D(64, 64, 1), nn.ReLU(),
                                  nn.Conv2D(64, 192, 3, padding=1), nn.ReLU(), nn.MaxPool2D(3, 2, 1))
        self.i3 = nn.Sequential(_Inception(192, 64, 96, 128, 16, 32, 32), _Inception(256, 128, 128, 192, 32, 96, 64), nn.MaxPool2D(3, 2, 1))
        self.i4a = _Inception(480, 192, 96, 208, 16, 48, 64)
        self.i4 = nn.Sequential(_Inception(512, 160, 112, 224, 24, 64, 64), _Inception(512, 128, 128, 256, 24, 64, 64), _Inception(512, 112, 144, 288, 32, 64, 64))
        self.i4e = nn.Sequential(_Inception(528, 256, 160, 320, 32, 128, 128), nn.MaxPool2D(3, 2, 1))
        self.i5 = nn.Sequential(_Inception(832, 256, 160, 320, 32, 128, 128), _Inception(832, 384, 192, 384, 48, 128, 128))
        if with_pool:
            self.pool = nn.AdaptiveAvgPool2D(1)
        if num_classes > 0:
            self.drop = nn.Dropout(0.4)
            self.fc = nn.Linear(1024, num_classes)
            self.aux1 = nn.Sequential(nn.AdaptiveAvgPool2D(4), nn.Conv2D(512, 128, 1), nn.ReLU(), nn.Flatten(), nn.Linear(2048, 1024), nn.ReLU(), nn.Dropout(0.7), nn.Linear(1024, num_classes))
            self.aux2 = nn.Sequential(nn.AdaptiveAvgPool2D(4), nn.Conv2D(528, 128, 1), nn.ReLU(), nn.Flatten(), nn.Linear(2048, 1024), nn.ReLU(), nn.Dropout(0.7), nn.Linear(1024, num_classes))

    def forward(self, x):
        x = self.i3(self.stem(x))
        a = self.i4a(x)
        b = self.i4(a)
        x = self.i5(self.i4e(b))
        if self.with_pool:
            x = self.pool(x)
        if self.num_classes > 0:
            out = self.fc(self.drop(x.flatten(1)))
            return out, self.aux1(a), self.aux2(b)
        return x


def googlenet(pretrained=False, **kw):
    _no_pretrained(pretrained)
    return GoogLeNet(**kw)


class _IncA(nn.Layer):
    def __init__(self, inp, pool):
        super().__init__()
        self.b1 = _cbr(inp, 64, 1)
        self.b5 = nn.Sequential(_cbr(inp, 48, 1), _cbr(48, 64, 5, p=2))
        self.b3 = nn.Sequential(_cbr(inp, 64, 1), _cbr(64, 96, 3, p=1), _cbr(96, 96, 3, p=1))
        self.bp = nn.Sequential(nn.AvgPool2D(3, 1, 1, exclusive=False), _cbr(inp, pool, 1))

    def forward(self, x):
        return torch.cat([self.b1(x), self.b5(x), self.b3(x), self.bp(x)], 1)


class _IncB(nn.Layer):
    def __init__(self, inp):
        super().__init__()
        self.b3 = _cbr(inp, 384, 3, 2)
        self.bd = nn.Sequential(_cbr(inp, 64, 1), _cbr(64, 96, 3, p=1), _cbr(96, 96, 3, 2))
        self.bp = nn.MaxPool2D(3, 2)

    def forward(self, x):
        return torch.cat([self.b3(x), self.bd(x), self.bp(x)], 1)


class _IncC(nn.Layer):
    def __init__(self, inp, c7):
        super().__init__()
        self.b1 = _cbr(inp, 192, 1)
        self.b7 = nn.Sequential(_cbr(inp, c7, 1), _cbr(c7, c7, (1, 7), p=(0, 3)), _cbr(c7, 192, (7, 1), p=(3, 0)))
        self.bd = nn.Sequential(_cbr(inp, c7, 1), _cbr(c7, c7, (7, 1), p=(3, 0)), _cbr(c7, c7, (1, 7), p=(0, 3)), _cbr(c7, c7, (7, 1), p=(3, 0)), _cbr(c7, 192, (1, 7), p=(0, 3)))
        self.bp = nn.Sequential(nn.AvgPool2D(3, 1, 1, exclusive=False), _cbr(inp, 192, 1))

    def forward(self, x):
        return torch.cat([self.b1(x), self.b7(x), self.bd(x), self.bp(x)], 1)


class _IncD(nn.Layer):
    def __init__(self, inp):
        super().__init__()
        self.b3 = nn.Sequential(_cbr(inp, 192, 1), _cbr(192, 320, 3, 2))
        self.b7 = nn.Sequential(_cbr(inp, 192, 1), _cbr(192, 192, (1, 7), p=(0, 3)), _cbr(192, 192, (7, 1), p=(3, 0)), _cbr(192, 192, 3, 2))
        self.bp = nn.MaxPool2D(3, 2)

    def forward(self, x):
        return torch.cat([self.b3(x), self.b7(x), self.bp(x)], 1)


class _IncE(nn.Layer):
    def __init__(self, inp):
        super().__init__()
        self.b1 = _cbr(inp, 320, 1)
        self.b3 = _cbr(inp, 384, 1)
        self.b3a, self.b3b = _cbr(384, 384, (1, 3), p=(0, 1)), _cbr(384, 384, (3, 1), p=(1, 0))
        self.bd = nn.Sequential(_cbr(inp, 448, 1), _cbr(448, 384, 3, p=1))
        self.bda, self.bdb = _cbr(384, 384, (1, 3), p=(0, 1)), _cbr(384, 384, (3, 1), p=(1, 0))
        self.bp = nn.Sequential(nn.AvgPool2D(3, 1, 1, exclusive=False), _cbr(inp, 192, 1))

    def forward(self, x):
        b3 = self.b3(x)
        bd = self.bd(x)
        return torch.cat([self.b1(x), self.b3a(b3), self.b3b(b3), self.bda(bd), self.bdb(bd), self.bp(x)], 1)


class InceptionV3(nn.Layer):
    def __init__(self, num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        self.stem = nn.Sequential(_cbr(3, 32, 3, 2), _cbr(32, 32, 3), _cbr(32, 64, 3, p=1), nn.MaxPool2D(3, 2), _cbr(64, 80, 1), _cbr(80, 192, 3), nn.MaxPool2D(3, 2))
        self.blocks = nn.Sequential(_IncA(192, 32), _IncA(256, 64), _IncA(288, 64), _IncB(288), _IncC(768, 128), _IncC(768, 160), _IncC(768, 160),
                                    _IncC(768, 192), _IncD(768), _IncE(1280), _IncE(2048))
        if with_pool:
            self.pool = nn.AdaptiveAvgPool2D(1)
        if num_classes > 0:
            self.dropout = nn.Dropout(0.2, mode="downscale_in_infer")
            self.fc = nn.Linear(2048, num_classes)

    def forward(self, x):
        x = self.blocks(self.stem(x))
        if self.with_pool:
            x = self.pool(x)
        if self.num_classes > 0:
            x = self.fc(self.dropout(x.flatten(1)))
        return x


def inception_v3(pretrained=False, **kw):
    _no_pretrained(pretrained)
    return InceptionV3(**kw)
