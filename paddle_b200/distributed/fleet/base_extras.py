"""fleet.Fleet / UtilBase / Role / data generators. Parity: fleet/fleet.py, base/util_factory.py, base/role_maker.py, data_generator/."""
from __future__ import annotations

import sys


class Role:
    WORKER = 1
    SERVER = 2
    HETER_WORKER = 3
    ALL = 4
    COORDINATOR = 5


class UtilBase:
    """Small collective helpers on python objects (all_reduce / barrier / all_gather over the default group)."""

    def all_reduce(self, input, mode="sum", comm_world="worker"):
        import numpy as np
        import torch
        import torch.distributed as dist

        t = torch.as_tensor(np.asarray(input, dtype=np.float64))
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[mode])
        return t.numpy()

    def barrier(self, comm_world="worker"):
        import torch.distributed as dist

        if dist.is_initialized():
            dist.barrier()

    def all_gather(self, input, comm_world="worker"):
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size() == 1:
            return [input]
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, input)
        return out

    def get_file_shard(self, files):
        from .. import env

        n, r = env.get_world_size(), env.get_rank()
        per, rem = divmod(len(files), n)
        lo = r * per + min(r, rem)
        return files[lo:lo + per + (1 if r < rem else 0)]

    def print_on_rank(self, message, rank_id):
        from .. import env

        if env.get_rank() == rank_id:
            print(message)


class Fleet:
    """Object form of the module-level fleet API (`fleet = Fleet()` in the reference)."""

    def __getattr__(self, name):
        from .. import fleet as _f

        return getattr(_f, name)

    @property
    def util(self):
        return UtilBase()


class MultiSlotDataGenerator:
    """Parity: fleet/data_generator/data_generator.py: turns user samples into the multi-slot text protocol of the PS data feed."""

    def __init__(self):
        self._line_limit, self.batch_size_ = None, 32

    def set_batch(self, batch_size):
        self.batch_size_ = batch_size

    def generate_sample(self, line):
        raise NotImplementedError("override generate_sample(line) -> generator of [(slot_name, [values]), ...]")

    def generate_batch(self, samples):
        def it():
            yield from samples
        return it

    def _gen_str(self, sample):
        parts = []
        for name, vals in sample:
            parts.append(str(len(vals)))
            parts.extend(str(v) for v in vals)
        return " ".join(parts) + "\n"

    def run_from_stdin(self):
        for line in sys.stdin:
            for sample in self.generate_sample(line)():
                sys.stdout.write(self._gen_str(sample))

    def run_from_memory(self):
        out = []
        for sample in self.generate_sample(None)():
            out.append(self._gen_str(sample))
        return out


class MultiSlotStringDataGenerator(MultiSlotDataGenerator):
    pass
