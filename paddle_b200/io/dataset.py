"""Datasets. Parity: python/paddle/io/dataloader/dataset.py."""
from __future__ import annotations

import bisect
import math

import numpy as np
import torch


class Dataset:
    def __getitem__(self, idx):
        raise NotImplementedError(f"'__getitem__' not implemented in {type(self).__name__}")

    def __len__(self):
        raise NotImplementedError(f"'__len__' not implemented in {type(self).__name__}")


class IterableDataset(Dataset):
    def __iter__(self):
        raise NotImplementedError

    def __getitem__(self, idx):
        raise RuntimeError("IterableDataset does not support __getitem__")

    def __len__(self):
        # TypeError (not RuntimeError) so that list(ds) / length_hint fall back to plain iteration
        raise TypeError("IterableDataset does not support __len__")


class TensorDataset(Dataset):
    def __init__(self, tensors):
        if not all(t.shape[0] == tensors[0].shape[0] for t in tensors):
            raise ValueError("tensors must share the first dimension")
        self.tensors = tensors

    def __getitem__(self, index):
        return tuple(t[index] for t in self.tensors)

    def __len__(self):
        return self.tensors[0].shape[0]


class ComposeDataset(Dataset):
    def __init__(self, datasets):
        self.datasets = list(datasets)
        n = len(self.datasets[0])
        assert all(len(d) == n for d in self.datasets), "lengths of datasets should be same"

    def __len__(self):
        return len(self.datasets[0])

    def __getitem__(self, idx):
        sample = []
        for d in self.datasets:
            s = d[idx]
            sample.extend(s if isinstance(s, (list, tuple)) else [s])
        return tuple(sample)


class ChainDataset(IterableDataset):
    def __init__(self, datasets):
        self.datasets = list(datasets)

    def __iter__(self):
        for d in self.datasets:
            yield from d


class ConcatDataset(Dataset):
    def __init__(self, datasets):
        self.datasets = list(datasets)
        self.cumulative_sizes = list(np.cumsum([len(d) for d in self.datasets]))

    def __len__(self):
        return self.cumulative_sizes[-1]

    def __getitem__(self, idx):
        if idx < 0:
            idx += len(self)
        di = bisect.bisect_right(self.cumulative_sizes, idx)
        return self.datasets[di][idx - (self.cumulative_sizes[di - 1] if di > 0 else 0)]


class Subset(Dataset):
    def __init__(self, dataset, indices):
        self.dataset, self.indices = dataset, list(indices)

    def __getitem__(self, idx):
        return self.dataset[self.indices[idx]]

    def __len__(self):
        return len(self.indices)


def random_split(dataset, lengths, generator=None):
    n = len(dataset)
    if math.isclose(sum(lengths), 1.0) and sum(lengths) <= 1.0 + 1e-6 and all(0 <= l <= 1 for l in lengths):
        counts = [int(math.floor(n * f)) for f in lengths]
        for i in range(n - sum(counts)):
            counts[i % len(counts)] += 1
        lengths = counts
    if sum(lengths) != n:
        raise ValueError("Sum of input lengths does not equal the length of the input dataset!")
    perm = np.random.permutation(n).tolist()
    out, off = [], 0
    for l in lengths:
        out.append(Subset(dataset, perm[off:off + l]))
        off += l
    return out
