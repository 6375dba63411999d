"""Recurrent layers. Parity: python/paddle/nn/layer/rnn.py (SimpleRNN/LSTM/GRU, cells, RNN, BiRNN).

Gate order follows paddle: LSTM i,f,g(c~),o ; GRU r,z,c. Weights: weight_ih [G*H, I], weight_hh [G*H, H].
"""
from __future__ import annotations

import math

import torch

from ..tensor import Tensor
from . import initializer as I
from .container import LayerList
from .layer import Layer


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


class RNNCellBase(Layer):
    def get_initial_states(self, batch_ref, shape=None, dtype=None, init_value=0.0, batch_dim_idx=0):
        b = batch_ref.shape[batch_dim_idx]
        shape = shape or self.state_shape
        mk = lambda s: _w(torch.full([b, *s], init_value, dtype=batch_ref.dtype, device=batch_ref.device))
        if isinstance(shape[0], (list, tuple)):
            return tuple(mk(s) for s in shape)
        return mk(shape)


class SimpleRNNCell(RNNCellBase):
    def __init__(self, input_size, hidden_size, activation="tanh", weight_ih_attr=None, weight_hh_attr=None, bias_ih_attr=None, bias_hh_attr=None, name=None):
        super().__init__()
        std = 1.0 / math.sqrt(hidden_size)
        u = I.Uniform(-std, std)
        self.weight_ih = self.create_parameter([hidden_size, input_size], weight_ih_attr, default_initializer=u)
        self.weight_hh = self.create_parameter([hidden_size, hidden_size], weight_hh_attr, default_initializer=u)
        self.bias_ih = self.create_parameter([hidden_size], bias_ih_attr, is_bias=True, default_initializer=u)
        self.bias_hh = self.create_parameter([hidden_size], bias_hh_attr, is_bias=True, default_initializer=u)
        self.input_size, self.hidden_size, self.activation = input_size, hidden_size, activation

    @property
    def state_shape(self):
        return (self.hidden_size,)

    def forward(self, inputs, states=None):
        if states is None:
            states = self.get_initial_states(inputs)
        z = inputs @ self.weight_ih.t() + states @ self.weight_hh.t()
        if self.bias_ih is not None:
            z = z + self.bias_ih
        if self.bias_hh is not None:
            z = z + self.bias_hh
        h = torch.tanh(z) if self.activation == "tanh" else torch.relu(z)
        return h, h


class LSTMCell(RNNCellBase):
    def __init__(self, input_size, hidden_size, weight_ih_attr=None, weight_hh_attr=None, bias_ih_attr=None, bias_hh_attr=None, proj_size=0, name=None):
        super().__init__()
        std = 1.0 / math.sqrt(hidden_size)
        u = I.Uniform(-std, std)
        self.weight_ih = self.create_parameter([4 * hidden_size, input_size], weight_ih_attr, default_initializer=u)
        self.weight_hh = self.create_parameter([4 * hidden_size, proj_size or hidden_size], weight_hh_attr, default_initializer=u)
        self.bias_ih = self.create_parameter([4 * hidden_size], bias_ih_attr, is_bias=True, default_initializer=u)
        self.bias_hh = self.create_parameter([4 * hidden_size], bias_hh_attr, is_bias=True, default_initializer=u)
        self.proj_size = proj_size
        if proj_size:
            self.weight_ho = self.create_parameter([hidden_size, proj_size], default_initializer=u)
        self.input_size, self.hidden_size = input_size, hidden_size

    @property
    def state_shape(self):
        return ((self.proj_size or self.hidden_size,), (self.hidden_size,))

    def forward(self, inputs, states=None):
        if states is None:
            states = self.get_initial_states(inputs)
        h, c = states
        gates = inputs @ self.weight_ih.t() + h @ self.weight_hh.t()
        if self.bias_ih is not None:
            gates = gates + self.bias_ih
        if self.bias_hh is not None:
            gates = gates + self.bias_hh
        i, f, g, o = torch.chunk(gates, 4, -1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        if self.proj_size:
            h = h @ self.weight_ho
        return h, (h, c)


class GRUCell(RNNCellBase):
    def __init__(self, input_size, hidden_size, weight_ih_attr=None, weight_hh_attr=None, bias_ih_attr=None, bias_hh_attr=None, name=None):
        super().__init__()
        std = 1.0 / math.sqrt(hidden_size)
        u = I.Uniform(-std, std)
        self.weight_ih = self.create_parameter([3 * hidden_size, input_size], weight_ih_attr, default_initializer=u)
        self.weight_hh = self.create_parameter([3 * hidden_size, hidden_size], weight_hh_attr, default_initializer=u)
        self.bias_ih = self.create_parameter([3 * hidden_size], bias_ih_attr, is_bias=True, default_initializer=u)
        self.bias_hh = self.create_parameter([3 * hidden_size], bias_hh_attr, is_bias=True, default_initializer=u)
        self.input_size, self.hidden_size = input_size, hidden_size

    @property
    def state_shape(self):
        return (self.hidden_size,)

    def forward(self, inputs, states=None):
        if states is None:
            states = self.get_initial_states(inputs)
        h = states
        xg = inputs @ self.weight_ih.t() + (self.bias_ih if self.bias_ih is not None else 0)
        hg = h @ self.weight_hh.t() + (self.bias_hh if self.bias_hh is not None else 0)
        xr, xz, xc = torch.chunk(xg, 3, -1)
        hr, hz, hc = torch.chunk(hg, 3, -1)
        r, z = torch.sigmoid(xr + hr), torch.sigmoid(xz + hz)
        c = torch.tanh(xc + r * hc)
        h = (h - c) * z + c
        return h, h


class RNN(Layer):
    """Wraps a cell into a sequence layer. Parity: nn/layer/rnn.py:RNN."""

    def __init__(self, cell, is_reverse=False, time_major=False):
        super().__init__()
        self.cell, self.is_reverse, self.time_major = cell, is_reverse, time_major

    def forward(self, inputs, initial_states=None, sequence_length=None, **kwargs):
        x = inputs if self.time_major else inputs.transpose([1, 0, 2])
        T_ = x.shape[0]
        states = initial_states if initial_states is not None else self.cell.get_initial_states(x, batch_dim_idx=1)
        outs = [None] * T_
        steps = range(T_ - 1, -1, -1) if self.is_reverse else range(T_)
        seq = None if sequence_length is None else sequence_length.as_subclass(torch.Tensor).to(x.device)
        for t in steps:
            out, new_states = self.cell(x[t], states, **kwargs)
            if seq is not None:
                m = (seq > t).to(out.dtype).unsqueeze(-1)
                out = out * m
                if isinstance(new_states, (tuple, list)):
                    new_states = tuple(ns * m + s * (1 - m) for ns, s in zip(new_states, states))
                else:
                    new_states = new_states * m + states * (1 - m)
            states = new_states
            outs[t] = out
        y = torch.stack(outs, 0)
        if not self.time_major:
            y = y.transpose([1, 0, 2])
        return y, states


class BiRNN(Layer):
    def __init__(self, cell_fw, cell_bw, time_major=False):
        super().__init__()
        self.rnn_fw = RNN(cell_fw, False, time_major)
        self.rnn_bw = RNN(cell_bw, True, time_major)

    def forward(self, inputs, initial_states=None, sequence_length=None, **kwargs):
        s_fw, s_bw = (None, None) if initial_states is None else initial_states
        y1, st1 = self.rnn_fw(inputs, s_fw, sequence_length, **kwargs)
        y2, st2 = self.rnn_bw(inputs, s_bw, sequence_length, **kwargs)
        return torch.cat([y1, y2], -1), (st1, st2)


class _RNNBase(Layer):
    _mode = "RNN_TANH"

    def __init__(self, input_size, hidden_size, num_layers=1, direction="forward", time_major=False, dropout=0.0,
                 weight_ih_attr=None, weight_hh_attr=None, bias_ih_attr=None, bias_hh_attr=None, activation="tanh", proj_size=0, name=None):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.time_major, self.dropout = time_major, dropout
        self.num_directions = 2 if direction in ("bidirect", "bidirectional") else 1
        self.proj_size = proj_size
        kw = dict(weight_ih_attr=weight_ih_attr, weight_hh_attr=weight_hh_attr, bias_ih_attr=bias_ih_attr, bias_hh_attr=bias_hh_attr)

        def mk(isz):
            if self._mode == "LSTM":
                return LSTMCell(isz, hidden_size, proj_size=proj_size, **kw)
            if self._mode == "GRU":
                return GRUCell(isz, hidden_size, **kw)
            return SimpleRNNCell(isz, hidden_size, activation, **kw)

        layers = []
        out_sz = proj_size or hidden_size
        for l in range(num_layers):
            isz = input_size if l == 0 else out_sz * self.num_directions
            if self.num_directions == 2:
                layers.append(BiRNN(mk(isz), mk(isz), time_major))
            else:
                layers.append(RNN(mk(isz), False, time_major))
        self.layers = LayerList(layers)

    def _split_states(self, initial_states):
        if initial_states is None:
            return [None] * self.num_layers
        nd = self.num_directions
        if self._mode == "LSTM":
            h0, c0 = initial_states
            per = []
            for l in range(self.num_layers):
                if nd == 2:
                    per.append(((h0[2 * l], c0[2 * l]), (h0[2 * l + 1], c0[2 * l + 1])))
                else:
                    per.append((h0[l], c0[l]))
            return per
        h0 = initial_states
        return [((h0[2 * l], h0[2 * l + 1]) if nd == 2 else h0[l]) for l in range(self.num_layers)]

    def forward(self, inputs, initial_states=None, sequence_length=None):
        from . import functional as F

        x = inputs
        finals = []
        for l, (layer, st) in enumerate(zip(self.layers, self._split_states(initial_states))):
            x, fs = layer(x, st, sequence_length)
            finals.append(fs)
            if self.dropout > 0 and l < self.num_layers - 1:
                x = F.dropout(x, self.dropout, training=self.training)
        flat = []
        for fs in finals:
            flat.extend(list(fs) if self.num_directions == 2 else [fs])
        if self._mode == "LSTM":
            h = torch.stack([s[0] for s in flat], 0)
            c = torch.stack([s[1] for s in flat], 0)
            return x, (h, c)
        return x, torch.stack(flat, 0)


class SimpleRNN(_RNNBase):
    _mode = "RNN_TANH"

    def __init__(self, input_size, hidden_size, num_layers=1, direction="forward", time_major=False, dropout=0.0, activation="tanh",
                 weight_ih_attr=None, weight_hh_attr=None, bias_ih_attr=None, bias_hh_attr=None, name=None):
        super().__init__(input_size, hidden_size, num_layers, direction, time_major, dropout, weight_ih_attr, weight_hh_attr, bias_ih_attr, bias_hh_attr,
                         activation=activation, name=name)


class LSTM(_RNNBase):
    _mode = "LSTM"

    def __init__(self, input_size, hidden_size, num_layers=1, direction="forward", time_major=False, dropout=0.0,
                 weight_ih_attr=None, weight_hh_attr=None, bias_ih_attr=None, bias_hh_attr=None, proj_size=0, name=None):
        super().__init__(input_size, hidden_size, num_layers, direction, time_major, dropout, weight_ih_attr, weight_hh_attr, bias_ih_attr, bias_hh_attr,
                         proj_size=proj_size, name=name)


class GRU(_RNNBase):
    _mode = "GRU"

    def __init__(self, input_size, hidden_size, num_layers=1, direction="forward", time_major=False, dropout=0.0,
                 weight_ih_attr=None, weight_hh_attr=None, bias_ih_attr=None, bias_hh_attr=None, name=None):
        super().__init__(input_size, hidden_size, num_layers, direction, time_major, dropout, weight_ih_attr, weight_hh_attr, bias_ih_attr, bias_hh_attr,
                         name=name)
