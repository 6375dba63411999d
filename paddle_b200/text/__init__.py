"""paddle.text. Parity: python/paddle/text/{viterbi_decode.py, datasets/*}."""
from __future__ import annotations

import os
import tarfile

import numpy as np
import torch

from ..io import Dataset
from ..nn.layer import Layer
from ..tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


def viterbi_decode(potentials, transition_params, lengths, include_bos_eos_tag=True, name=None):
    """Parity: python/paddle/text/viterbi_decode.py (viterbi_decode kernel)."""
    pot, trans, lens = _raw(potentials).float(), _raw(transition_params).float(), _raw(lengths).long()
    B, T, N = pot.shape
    if include_bos_eos_tag:
        start, stop = trans[-2], trans[:, -1]   # BOS row = N-2, EOS column = N-1 (reference layout)
        alpha = pot[:, 0] + start[None]
    else:
        alpha = pot[:, 0].clone()
    hist = []
    for t in range(1, T):
        score = alpha[:, :, None] + trans[None]
        best, idx = score.max(1)
        new = best + pot[:, t]
        mask = (t < lens)[:, None]
        alpha = torch.where(mask, new, alpha)
        hist.append(torch.where(mask, idx, torch.arange(N, device=pot.device)[None].expand(B, N)))
    if include_bos_eos_tag:
        alpha = alpha + stop[None]
    scores, last = alpha.max(1)
    path = [last]
    for idx in reversed(hist):
        last = torch.gather(idx, 1, last[:, None]).squeeze(1)
        path.append(last)
    path = torch.stack(path[::-1], 1)
    pos = torch.arange(T, device=pot.device)[None]
    path = torch.where(pos < lens[:, None], path, torch.zeros_like(path))
    return scores.as_subclass(Tensor), path.as_subclass(Tensor)


class ViterbiDecoder(Layer):
    def __init__(self, transitions, include_bos_eos_tag=True, name=None):
        super().__init__()
        self.transitions, self.include_bos_eos_tag = transitions, include_bos_eos_tag

    def forward(self, potentials, lengths):
        return viterbi_decode(potentials, self.transitions, lengths, self.include_bos_eos_tag)


def _need(path, what):
    if path is None or not os.path.exists(path):
        raise RuntimeError(f"{what}: data_file '{path}' not found and downloading is disabled (no network)")


class UCIHousing(Dataset):
    def __init__(self, data_file=None, mode="train", download=True):
        _need(data_file, "UCIHousing")
        data = np.fromfile(data_file, sep=" ").reshape(-1, 14)
        mx, mn, avg = data.max(0), data.min(0), data.mean(0)
        for i in range(13):
            data[:, i] = (data[:, i] - avg[i]) / (mx[i] - mn[i])
        off = int(data.shape[0] * 0.8)
        self.data = (data[:off] if mode == "train" else data[off:]).astype("float32")

    def __getitem__(self, idx):
        return self.data[idx, :-1], self.data[idx, -1:]

    def __len__(self):
        return len(self.data)


class Imdb(Dataset):
    def __init__(self, data_file=None, mode="train", cutoff=150, download=True):
        import re
        import string

        _need(data_file, "Imdb")
        self.docs, self.labels = [], []
        pat = re.compile(rf"aclImdb/{mode}/(pos|neg)/.*\.txt$")
        freq = {}
        texts = []
        with tarfile.open(data_file) as tf:
            for m in tf:
                mt = pat.match(m.name)
                if mt:
                    words = tf.extractfile(m).read().decode("latin-1").lower().translate(str.maketrans("", "", string.punctuation)).split()
                    texts.append((words, 0 if mt.group(1) == "pos" else 1))
                    for w in words:
                        freq[w] = freq.get(w, 0) + 1
        vocab = sorted([w for w, c in freq.items() if c > cutoff], key=lambda w: (-freq[w], w))
        self.word_idx = {w: i for i, w in enumerate(vocab)}
        self.word_idx["<unk>"] = len(vocab)
        unk = self.word_idx["<unk>"]
        for words, lab in texts:
            self.docs.append(np.array([self.word_idx.get(w, unk) for w in words]))
            self.labels.append(np.array([lab]))

    def __getitem__(self, idx):
        return self.docs[idx], self.labels[idx]

    def __len__(self):
        return len(self.docs)


class Imikolov(Dataset):
    def __init__(self, data_file=None, data_type="NGRAM", window_size=-1, mode="train", min_word_freq=50, download=True):
        _need(data_file, "Imikolov")
        name = {"train": "./simple-examples/data/ptb.train.txt", "test": "./simple-examples/data/ptb.valid.txt"}[mode]
        with tarfile.open(data_file) as tf:
            train = tf.extractfile("./simple-examples/data/ptb.train.txt").read().decode().splitlines()
            lines = tf.extractfile(name).read().decode().splitlines()
        freq = {}
        for l in train:
            for w in l.split():
                freq[w] = freq.get(w, 0) + 1
        vocab = sorted([w for w, c in freq.items() if c > min_word_freq], key=lambda w: (-freq[w], w))
        self.word_idx = {w: i for i, w in enumerate(vocab)}
        for tok in ("<unk>", "<s>", "<e>"):
            self.word_idx.setdefault(tok, len(self.word_idx))
        unk = self.word_idx["<unk>"]
        self.data = []
        for l in lines:
            ids = [self.word_idx["<s>"]] + [self.word_idx.get(w, unk) for w in l.split()] + [self.word_idx["<e>"]]
            if data_type == "NGRAM":
                for i in range(window_size, len(ids) + 1):
                    self.data.append(tuple(ids[i - window_size:i]))
            else:
                self.data.append((ids[:-1], ids[1:]))

    def __getitem__(self, idx):
        return tuple(np.array(d) for d in self.data[idx])

    def __len__(self):
        return len(self.data)


class _LocalOnly(Dataset):
    _name = "dataset"

    def __init__(self, data_file=None, *a, **k):
        _need(data_file, self._name)
        raise NotImplementedError(f"{self._name}: local archive parsing is not implemented in this build")


class Movielens(_LocalOnly):
    _name = "Movielens"


class Conll05st(_LocalOnly):
    _name = "Conll05st"


class WMT14(_LocalOnly):
    _name = "WMT14"


class WMT16(_LocalOnly):
    _name = "WMT16"


__all__ = ["Conll05st", "Imdb", "Imikolov", "Movielens", "UCIHousing", "WMT14", "WMT16", "ViterbiDecoder", "viterbi_decode"]
