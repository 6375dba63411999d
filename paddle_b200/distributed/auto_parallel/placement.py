"""Placements. Parity: paddle/phi/core/distributed/auto_parallel/placement_types.h, python/.../placement_type.py."""


class Placement:
    def is_shard(self, dim=None):
        return False

    def is_replicated(self):
        return False

    def is_partial(self):
        return False

    def __ne__(self, o):
        return not self == o


class Replicate(Placement):
    def is_replicated(self):
        return True

    def __eq__(self, o):
        return isinstance(o, Replicate)

    def __hash__(self):
        return hash("R")

    def __repr__(self):
        return "Replicate()"


class Shard(Placement):
    def __init__(self, dim):
        self.dim = int(dim)

    def get_dim(self):
        return self.dim

    def is_shard(self, dim=None):
        return dim is None or dim == self.dim

    def __eq__(self, o):
        return isinstance(o, Shard) and o.dim == self.dim

    def __hash__(self):
        return hash(("S", self.dim))

    def __repr__(self):
        return f"Shard(dim={self.dim})"


class Partial(Placement):
    def __init__(self, reduce_type="sum"):
        rt = getattr(reduce_type, "name", reduce_type)
        rt = str(rt).lower().replace("reduceop.", "").replace("k_red_", "")
        self.reduce_type = rt[4:] if rt.startswith("kred") else rt          # ReduceType.kRedSum -> "sum"

    def is_partial(self):
        return True

    def __eq__(self, o):
        return isinstance(o, Partial) and o.reduce_type == self.reduce_type

    def __hash__(self):
        return hash(("P", self.reduce_type))

    def __repr__(self):
        return f"Partial(reduce_type={self.reduce_type})"
