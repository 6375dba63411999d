"""paddle.io. Parity: python/paddle/io/__init__.py."""
from .dataloader import DataLoader, default_collate_fn, default_convert_fn, get_worker_info  # noqa: F401
from .dataset import (ChainDataset, ComposeDataset, ConcatDataset, Dataset, IterableDataset, Subset, TensorDataset,  # noqa: F401
                      random_split)
from .sampler import (BatchSampler, DistributedBatchSampler, RandomSampler, Sampler, SequenceSampler, SubsetRandomSampler,  # noqa: F401
                      WeightedRandomSampler)
