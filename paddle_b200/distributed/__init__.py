"""paddle.distributed. Parity: python/paddle/distributed/__init__.py."""
from . import env, collective  # noqa: F401
from .collective import (Group, P2POp, ReduceOp, all_gather, all_gather_into_tensor, all_gather_object, all_reduce, alltoall,  # noqa: F401
                         alltoall_single, barrier, batch_isend_irecv, broadcast, broadcast_object_list, gather, get_backend, get_group,
                         irecv, is_available, isend, new_group, recv, reduce, reduce_scatter, reduce_scatter_tensor, scatter,
                         scatter_object_list, send, split, wait)
from .env import (ParallelEnv, destroy_process_group, get_rank, get_world_size, init_parallel_env, is_initialized)  # noqa: F401
from . import fleet  # noqa: F401


def __getattr__(name):
    import importlib

    lazy = {"DataParallel": ".data_parallel", "spawn": ".spawn", "launch": ".launch", "sharding": ".sharding", "checkpoint": ".checkpoint",
            "rpc": ".rpc", "auto_parallel": ".auto_parallel", "stream": ".stream", "utils": ".dist_utils", "ps": ".ps",
            "watchdog": ".watchdog", "auto_tuner": ".auto_tuner"}
    if name in lazy:
        mod = importlib.import_module(lazy[name], __name__)
        return getattr(mod, name) if name in ("DataParallel", "spawn") else mod
    ap = {"ProcessMesh", "shard_tensor", "dtensor_from_fn", "reshard", "Shard", "Replicate", "Partial", "shard_layer", "shard_optimizer",
          "shard_dataloader", "to_static", "Strategy", "DistModel", "Placement", "unshard_dtensor", "shard_scaler", "ShardingStage1",
          "ShardingStage2", "ShardingStage3", "DistAttr", "parallelize", "ColWiseParallel", "RowWiseParallel", "SequenceParallelBegin",
          "SequenceParallelEnd", "SequenceParallelEnable", "SequenceParallelDisable", "PrepareLayerInput", "PrepareLayerOutput", "SplitPoint"}
    if name in ap:
        mod = importlib.import_module(".auto_parallel", __name__)
        return getattr(mod, name)
    if name in ("save_state_dict", "load_state_dict"):
        mod = importlib.import_module(".checkpoint", __name__)
        return getattr(mod, name)
    if name in ("gloo_init_parallel_env", "gloo_barrier", "gloo_release", "QueueDataset", "InMemoryDataset", "CountFilterEntry", "ShowClickEntry",
                "ProbabilityEntry", "ParallelMode", "ReduceType", "io"):
        mod = importlib.import_module(".extras", __name__)
        return getattr(mod, name)
    if name in ("group_sharded_parallel", "save_group_sharded_model"):
        mod = importlib.import_module(".sharding", __name__)
        return getattr(mod, name)
    raise AttributeError(f"module 'paddle_b200.distributed' has no attribute '{name}'")

from .auto_parallel import Engine, LocalLayer, ShardDataloader, enable_auto_dp, to_distributed  # noqa: F401,E402
