"""Beam search decoding. Parity: python/paddle/nn/decode.py (BeamSearchDecoder, dynamic_decode)."""
from __future__ import annotations

import torch

from ..tensor import Tensor
from . import functional as F


def _w(t):
    return t.as_subclass(Tensor) if isinstance(t, torch.Tensor) and not isinstance(t, Tensor) else t


class BeamSearchDecoder:
    def __init__(self, cell, start_token, end_token, beam_size, embedding_fn=None, output_fn=None):
        self.cell, self.start_token, self.end_token, self.beam_size = cell, start_token, end_token, beam_size
        self.embedding_fn, self.output_fn = embedding_fn, output_fn

    @staticmethod
    def tile_beam_merge_with_batch(x, beam_size):
        x = x.as_subclass(torch.Tensor)
        return _w(x.unsqueeze(1).expand(x.shape[0], beam_size, *x.shape[1:]).reshape(-1, *x.shape[1:]))

    def _map(self, fn, s):
        if isinstance(s, (tuple, list)):
            return type(s)(self._map(fn, i) for i in s)
        return fn(s)

    def initialize(self, initial_cell_states):
        ref = initial_cell_states
        while isinstance(ref, (tuple, list)):
            ref = ref[0]
        b = ref.shape[0]
        self.batch = b
        dev = ref.device
        states = self._map(lambda s: self.tile_beam_merge_with_batch(s, self.beam_size), initial_cell_states)
        log_probs = torch.full((b, self.beam_size), -1e9, device=dev)
        log_probs[:, 0] = 0.0
        finished = torch.zeros(b, self.beam_size, dtype=torch.bool, device=dev)
        tokens = torch.full((b * self.beam_size,), self.start_token, dtype=torch.int64, device=dev)
        inputs = self.embedding_fn(_w(tokens)) if self.embedding_fn else _w(tokens)
        return inputs, (states, log_probs, finished), finished

    def step(self, time, inputs, states):
        cell_states, log_probs, finished = states
        out, new_cell = self.cell(inputs, cell_states)
        if self.output_fn is not None:
            out = self.output_fn(out)
        logp = torch.log_softmax(out.as_subclass(torch.Tensor).float(), -1)
        V = logp.shape[-1]
        logp = logp.reshape(self.batch, self.beam_size, V)
        # finished beams only extend with end_token at zero cost
        fin_mask = torch.full((V,), -1e9, device=logp.device)
        fin_mask[self.end_token] = 0.0
        logp = torch.where(finished.unsqueeze(-1), fin_mask, logp)
        total = (log_probs.unsqueeze(-1) + logp).reshape(self.batch, -1)
        top, idx = total.topk(self.beam_size, -1)
        parent = idx // V
        token = idx % V
        gather = (parent + torch.arange(self.batch, device=idx.device).unsqueeze(1) * self.beam_size).reshape(-1)
        new_cell = self._map(lambda s: _w(s.as_subclass(torch.Tensor)[gather]), new_cell)
        new_finished = torch.gather(finished, 1, parent) | (token == self.end_token)
        next_inputs = self.embedding_fn(_w(token.reshape(-1))) if self.embedding_fn else _w(token.reshape(-1))
        return (token, parent), (new_cell, top, new_finished), next_inputs, new_finished

    def finalize(self, outputs, final_states, sequence_lengths):
        tokens = torch.stack([o[0] for o in outputs], 0)
        parents = torch.stack([o[1] for o in outputs], 0)
        return F.gather_tree(_w(tokens), _w(parents)), final_states


def dynamic_decode(decoder, inits=None, max_step_num=None, output_time_major=False, impute_finished=False, is_test=False,
                   return_length=False, **kwargs):
    inputs, states, finished = decoder.initialize(inits)
    outputs = []
    step = 0
    lengths = torch.zeros_like(finished, dtype=torch.int64)
    while True:
        out, states, inputs, finished_new = decoder.step(step, inputs, states, **kwargs)
        lengths = lengths + (~finished).long()
        finished = finished_new
        outputs.append(out)
        step += 1
        if bool(finished.all()) or (max_step_num is not None and step >= max_step_num):
            break
    ids, final_states = decoder.finalize(outputs, states, lengths)
    if not output_time_major:
        ids = ids.transpose([1, 0, 2])
    return (ids, final_states, _w(lengths)) if return_length else (ids, final_states)
