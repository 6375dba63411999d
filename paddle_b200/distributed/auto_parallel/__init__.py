"""Semi-automatic parallelism (DistTensor API). Parity: python/paddle/distributed/auto_parallel/{api.py,process_mesh.py,
placement_type.py,intermediate/*.py} and the SPMD rules under paddle/phi/infermeta/spmd_rules/.

Design: a `DistTensor` is a `Tensor` subclass whose storage is the LOCAL shard and whose metadata (`process_mesh`,
`placements`, global shape) describes the global view. Ops on DistTensors go through `__torch_function__`, which applies
a sharding-propagation rule (elementwise / matmul / reductions / embedding) or falls back to "replicate then compute";
`reshard` implements the placement transitions with autograd-aware collectives. `parallelize` (the intermediate API)
swaps sublayers for the tensor-parallel fleet layers, which use the fused tcgen05-GEMM + peer-memory collective paths."""
from .process_mesh import ProcessMesh, get_mesh, set_mesh  # noqa: F401
from .placement import Partial, Placement, Replicate, Shard  # noqa: F401
from .api import (DistAttr, DistModel, DistTensor, ShardingStage1, ShardingStage2, ShardingStage3, Strategy, dtensor_from_fn,  # noqa: F401
                  reshard, shard_dataloader, shard_layer, shard_optimizer, shard_scaler, shard_tensor, to_static, unshard_dtensor)
from .intermediate import (ColWiseParallel, PrepareLayerInput, PrepareLayerOutput, RowWiseParallel, SequenceParallelBegin,  # noqa: F401
                           SequenceParallelDisable, SequenceParallelEnable, SequenceParallelEnd, SplitPoint, parallelize)
from .api import _ShardDataLoader as ShardDataloader  # noqa: F401,E402
from .engine import Engine, LocalLayer, enable_auto_dp, in_auto_dp_mode, to_distributed  # noqa: F401,E402
