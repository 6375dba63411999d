"""Hybrid-parallel topology. Parity: python/paddle/distributed/fleet/base/topology.py
(CommunicateTopology, HybridCommunicateGroup, ParallelMode)."""
from __future__ import annotations

import itertools
from functools import reduce

import numpy as np

from .. import collective as C
from .. import env

_HYBRID_PARALLEL_GROUP = None


class ParallelMode:
    DATA_PARALLEL = 0
    TENSOR_PARALLEL = 1
    PIPELINE_PARALLEL = 2
    SHARDING_PARALLEL = 3
    SEGMENT_PARALLEL = 4


class CommunicateTopology:
    def __init__(self, hybrid_group_names=("data", "pipe", "sharding", "sep", "model"), dims=(1, 1, 1, 1, 1)):
        self._parallel_names = list(hybrid_group_names)
        self._dims = list(dims)
        self._world_size = reduce(lambda a, b: a * b, self._dims, 1)
        ranges = [range(d) for d in self._dims]
        self._coords = list(itertools.product(*ranges))
        self._coord2rank = {c: i for i, c in enumerate(self._coords)}
        self._rank2coord = {i: c for i, c in enumerate(self._coords)}

    def get_hybrid_group_names(self):
        return self._parallel_names

    def get_dim(self, axis_name):
        return self._dims[self._parallel_names.index(axis_name)]

    get_dim_size = get_dim

    def world_size(self):
        return self._world_size

    def get_rank(self, **args):
        assert len(args) == len(self._dims)
        return self._coord2rank[tuple(args[n] for n in self._parallel_names)]

    def get_coord(self, rank):
        return self._rank2coord[rank]

    def get_axis_list(self, axis_name, index):
        ax = self._parallel_names.index(axis_name)
        return sorted(r for r, c in self._rank2coord.items() if c[ax] == index)

    def get_comm_list(self, axis_name):
        """All groups along `axis_name`: list of rank lists."""
        ax = self._parallel_names.index(axis_name)
        other = [range(d) for i, d in enumerate(self._dims) if i != ax]
        out = []
        for oc in itertools.product(*other):
            ranks = []
            for k in range(self._dims[ax]):
                c = list(oc)
                c.insert(ax, k)
                ranks.append(self._coord2rank[tuple(c)])
            out.append(ranks)
        return out

    def get_fused_ranks(self, fused_axis):
        axes = [self._parallel_names.index(a) for a in fused_axis]
        other_axes = [i for i in range(len(self._dims)) if i not in axes]
        out = []
        for oc in itertools.product(*[range(self._dims[i]) for i in other_axes]):
            ranks = []
            for fc in itertools.product(*[range(self._dims[i]) for i in axes]):
                c = [0] * len(self._dims)
                for i, v in zip(other_axes, oc):
                    c[i] = v
                for i, v in zip(axes, fc):
                    c[i] = v
                ranks.append(self._coord2rank[tuple(c)])
            out.append(sorted(ranks))
        return out

    def get_rank_from_stage(self, global_rank, **kwargs):
        c = list(self.get_coord(global_rank))
        for k, v in kwargs.items():
            c[self._parallel_names.index(k)] = v
        return self._coord2rank[tuple(c)]


class HybridCommunicateGroup:
    def __init__(self, topology: CommunicateTopology):
        self._topo = topology
        self.global_rank = env.get_rank()
        self.nranks = env.get_world_size()
        assert self.nranks == topology.world_size(), f"world size {self.nranks} != topology {topology.world_size()}"
        names = topology.get_hybrid_group_names()
        self._dp_degree = topology.get_dim("data")
        self._mp_degree = topology.get_dim("model")
        self._pp_degree = topology.get_dim("pipe")
        self._sharding_degree = topology.get_dim("sharding")
        self._sep_degree = topology.get_dim("sep") if "sep" in names else 1
        coord = topology.get_coord(self.global_rank)
        self._coord = dict(zip(names, coord))
        self._groups = {}
        for axis in names:
            self._groups[axis] = self._make_group(topology.get_comm_list(axis))
        # fused groups used by grad clip / param sync
        self._check_group = self._make_group(topology.get_fused_ranks([a for a in ("pipe", "sharding", "model") if a in names]))
        self._pp_mp_group = self._make_group(topology.get_fused_ranks(["pipe", "model"]))
        self._dp_sep_group = self._make_group(topology.get_fused_ranks(["data", "sep"])) if "sep" in names else self._groups["data"]
        self._dp_sharding_group = self._make_group(topology.get_fused_ranks(["data", "sharding"]))
        # pipeline neighbours
        stage = self._coord["pipe"]
        self.stage_id = stage
        self._pp_ranks = self._groups["pipe"][1]
        self.next_rank = self._pp_ranks[(stage + 1) % self._pp_degree]
        self.prev_rank = self._pp_ranks[(stage - 1) % self._pp_degree]
        global _HYBRID_PARALLEL_GROUP
        _HYBRID_PARALLEL_GROUP = self

    def _make_group(self, comm_lists):
        mine = None
        for ranks in comm_lists:
            if len(ranks) == 1 or self.nranks == 1:
                if self.global_rank in ranks:
                    mine = (C.Group(0, -1, ranks, None), ranks)
                continue
            g = C.new_group(ranks)
            if self.global_rank in ranks:
                mine = (g, ranks)
        return mine

    # ---- info ----------------------------------------------------------------
    def get_parallel_mode(self):
        if self._mp_degree == 1 and self._pp_degree == 1 and self._sharding_degree == 1 and self._sep_degree == 1:
            return ParallelMode.DATA_PARALLEL
        if self._pp_degree > 1:
            return ParallelMode.PIPELINE_PARALLEL
        if self._mp_degree > 1:
            return ParallelMode.TENSOR_PARALLEL
        if self._sharding_degree > 1:
            return ParallelMode.SHARDING_PARALLEL
        return ParallelMode.SEGMENT_PARALLEL

    def topology(self):
        return self._topo

    def get_global_rank(self):
        return self.global_rank

    # data parallel
    def get_data_parallel_rank(self):
        return self._coord["data"]

    def get_data_parallel_world_size(self):
        return self._dp_degree

    def get_data_parallel_group(self):
        return self._groups["data"][0]

    def get_data_parallel_group_src_rank(self):
        return self._groups["data"][1][0]

    # model parallel
    def get_model_parallel_rank(self):
        return self._coord["model"]

    def get_model_parallel_world_size(self):
        return self._mp_degree

    def get_model_parallel_group(self):
        return self._groups["model"][0]

    def get_model_parallel_group_src_rank(self):
        return self._groups["model"][1][0]

    # pipeline
    def get_stage_id(self):
        return self.stage_id

    def get_pipe_parallel_world_size(self):
        return self._pp_degree

    def get_pipe_parallel_group(self):
        return self._groups["pipe"][0]

    def is_first_stage(self):
        return self.stage_id == 0

    def is_last_stage(self):
        return self.stage_id == self._pp_degree - 1

    def get_p2p_groups(self):
        return None, None, None, None

    # sharding
    def get_sharding_parallel_rank(self):
        return self._coord["sharding"]

    def get_sharding_parallel_world_size(self):
        return self._sharding_degree

    def get_sharding_parallel_group(self):
        return self._groups["sharding"][0]

    def get_sharding_parallel_group_src_rank(self):
        return self._groups["sharding"][1][0]

    # sep
    def get_sep_parallel_rank(self):
        return self._coord.get("sep", 0)

    def get_sep_parallel_world_size(self):
        return self._sep_degree

    def get_sep_parallel_group(self):
        return self._groups["sep"][0] if "sep" in self._groups else None

    def get_sep_parallel_group_src_rank(self):
        g = self.get_sep_parallel_group()
        return g.ranks[0] if g is not None else self.global_rank

    def create_fuse_group(self, fused_strategy_list):
        """Communicator over the product of several axes, e.g. ["data", "sharding"] (every process must call it). Parity: topology.py."""
        from ..collective import new_group

        assert len(fused_strategy_list) > 0
        lists = self._topo.get_fused_ranks(fused_strategy_list) if hasattr(self._topo, "get_fused_ranks") else None
        if lists is None:
            import itertools

            names, dims = self._topo._parallel_names, self._topo._dims
            fused = [names.index(n) for n in fused_strategy_list]
            others = [i for i in range(len(names)) if i not in fused]
            lists = []
            for oc in itertools.product(*[range(dims[i]) for i in others]):
                ranks = []
                for fc in itertools.product(*[range(dims[i]) for i in fused]):
                    coord = [0] * len(names)
                    for i, c in zip(others, oc):
                        coord[i] = c
                    for i, c in zip(fused, fc):
                        coord[i] = c
                    ranks.append(self._topo.get_rank(**{names[i]: coord[i] for i in range(len(names))}))
                lists.append(sorted(ranks))
        mine = None
        for ranks in lists:
            g = new_group(ranks)
            if self.global_rank in ranks:
                mine = g
        return mine

    def get_check_parallel_group(self, sharding=False):
        return self._check_group[0]

    def get_pp_mp_parallel_group(self):
        return self._pp_mp_group[0]

    def get_dp_sep_parallel_group(self):
        return self._dp_sep_group[0]

    def get_dp_sharding_parallel_group(self):
        return self._dp_sharding_group[0]

    def get_rank_from_stage(self, stage_id, **kwargs):
        return self._topo.get_rank_from_stage(self.global_rank, pipe=stage_id, **kwargs)


def get_hybrid_communicate_group():
    return _HYBRID_PARALLEL_GROUP


def _set_hcg(h):
    global _HYBRID_PARALLEL_GROUP
    _HYBRID_PARALLEL_GROUP = h
