"""MNIST MLP (BASELINE config 1: CPU plumbing, .pdparams save/load). Parity: the hapi / quick-start MNIST examples."""
from .. import nn


class MnistMLP(nn.Layer):
    def __init__(self, hidden=(512, 256), num_classes=10):
        super().__init__()
        dims = [784, *hidden]
        layers = []
        for a, b in zip(dims[:-1], dims[1:]):
            layers += [nn.Linear(a, b), nn.ReLU()]
        self.net = nn.Sequential(nn.Flatten(), *layers, nn.Linear(dims[-1], num_classes))

    def forward(self, x):
        return self.net(x)
