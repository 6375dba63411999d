"""paddle.distributed.fleet. Parity: python/paddle/distributed/fleet/__init__.py, fleet.py, model.py, optimizer.py."""
from __future__ import annotations

from .. import env as _env
from . import mp_layers as meta_parallel_layers  # noqa: F401
from . import topology as _topo
from .mp_layers import (ColumnParallelLinear, ColumnSequenceParallelLinear, ParallelCrossEntropy, RowParallelLinear,  # noqa: F401
                        RowSequenceParallelLinear, VocabParallelEmbedding)
from .random import get_rng_state_tracker, model_parallel_random_seed  # noqa: F401
from .recompute import recompute, recompute_hybrid, recompute_sequential  # noqa: F401
from .strategy import DistributedStrategy  # noqa: F401
from .topology import CommunicateTopology, HybridCommunicateGroup, ParallelMode, get_hybrid_communicate_group  # noqa: F401

_state = {"strategy": None, "hcg": None, "initialized": False}


class _RoleMaker:
    def __init__(self, is_collective=True, **kw):
        self._is_collective = is_collective

    def _worker_index(self):
        return _env.get_rank()

    def _worker_num(self):
        return _env.get_world_size()


PaddleCloudRoleMaker = _RoleMaker
UserDefinedRoleMaker = _RoleMaker


def init(role_maker=None, is_collective=False, strategy=None, log_level="INFO"):
    """fleet.init: builds the hybrid topology (dp/pp/sharding/sep/mp) and its process groups."""
    strategy = strategy or DistributedStrategy()
    _state["strategy"] = strategy
    from . import ps_mode

    if not is_collective and ps_mode.ps_env_present():      # parameter-server job (TRAINING_ROLE / PADDLE_PSERVERS_IP_PORT_LIST)
        _state["ps"] = ps_mode.init(strategy)
        _state["initialized"] = True
        return None
    _state["ps"] = None
    if not _env.is_initialized():
        _env.init_parallel_env()
    hc = strategy.hybrid_configs
    world = _env.get_world_size()
    mp, pp, sh, sep = hc["mp_degree"], hc["pp_degree"], hc["sharding_degree"], hc.get("sep_degree", 1) or 1
    dp = hc["dp_degree"]
    if dp * mp * pp * sh * sep != world:
        dp = world // (mp * pp * sh * sep)
        hc["dp_degree"] = dp
    assert dp * mp * pp * sh * sep == world, f"hybrid degrees {dp}x{pp}x{sh}x{sep}x{mp} do not match world size {world}"
    names_map = {"dp": "data", "pp": "pipe", "sharding": "sharding", "sep": "sep", "mp": "model"}
    order = hc.get("order") or ["dp", "pp", "sharding", "sep", "mp"]
    dims_map = {"dp": dp, "pp": pp, "sharding": sh, "sep": sep, "mp": mp}
    topo = CommunicateTopology([names_map[o] for o in order], [dims_map[o] for o in order])
    hcg = HybridCommunicateGroup(topo)
    _state["hcg"] = hcg
    _state["initialized"] = True
    seed = strategy.tensor_parallel_configs.get("tensor_init_seed", -1)
    model_parallel_random_seed(seed if seed and seed > 0 else None)
    return None


def _ps():
    return _state.get("ps")


def is_first_worker():
    return (_ps().index == 0 and not _ps().is_server) if _ps() else _env.get_rank() == 0


def worker_index():
    return _ps().index if _ps() else _env.get_rank()


def worker_num():
    return _ps().n_workers if _ps() else _env.get_world_size()


def server_num():
    return _ps().n_servers if _ps() else 0


def server_index():
    return _ps().index if (_ps() and _ps().is_server) else -1


def server_endpoints(to_string=False):
    eps = list(_ps().endpoints) if _ps() else []
    return ",".join(eps) if to_string else eps


def is_worker():
    return not _ps().is_server if _ps() else True


def is_server():
    return _ps().is_server if _ps() else False


def barrier_worker():
    if _ps():
        return
    from .. import collective

    collective.barrier()


def init_server(*args, **kwargs):
    from . import ps_mode

    return ps_mode.init_server(*args, **kwargs)


def run_server():
    from . import ps_mode

    return ps_mode.run_server()


def init_worker(scopes=None):
    if _ps():
        from . import ps_mode

        return ps_mode.init_worker()


def stop_worker():
    if _ps():
        from . import ps_mode

        ps_mode.stop_worker()


def save_persistables(executor=None, dirname=None, main_program=None, mode=0):
    """PS mode: every server writes its tables to `dirname`; collective mode: plain static persistables save."""
    if _ps():
        from . import ps_mode

        return ps_mode.client().save(dirname)
    from .. import extras

    return extras.io.save_persistables(executor, dirname, main_program)


def get_strategy():
    return _state["strategy"]


def distributed_model(model):
    """Wraps `model` according to the hybrid topology. Parity: fleet/model.py:distributed_model."""
    from . import hybrid

    return hybrid.distributed_model(model, _state["hcg"], _state["strategy"] or DistributedStrategy())


def distributed_optimizer(optimizer, strategy=None):
    from . import hybrid

    if strategy is not None:
        _state["strategy"] = strategy
    return hybrid.HybridParallelOptimizer(optimizer, _state["hcg"], _state["strategy"] or DistributedStrategy())


def distributed_scaler(scaler):
    from . import hybrid

    return hybrid.distributed_scaler(scaler, _state["hcg"])


class _MetaParallel:
    """fleet.meta_parallel namespace."""

    def __getattr__(self, name):
        from . import hybrid, mp_layers, pipeline

        for mod in (mp_layers, pipeline, hybrid):
            if hasattr(mod, name):
                return getattr(mod, name)
        if name in ("get_rng_state_tracker", "model_parallel_random_seed"):
            from . import random as r

            return getattr(r, name)
        raise AttributeError(name)


meta_parallel = _MetaParallel()

from .base_extras import Fleet, MultiSlotDataGenerator, MultiSlotStringDataGenerator, Role, UtilBase  # noqa: F401,E402
from . import utils  # noqa: F401,E402


def __getattr__(name):
    if name == "auto":      # paddle.distributed.fleet.auto
        import importlib

        return importlib.import_module(__name__ + ".auto")
    raise AttributeError(name)
