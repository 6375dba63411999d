"""Every module path a reference user imports from resolves (paddle_b200/_compat_paths.py), and `install_as_paddle()` serves
`import paddle.x.y` with the very same module objects."""
import importlib
import subprocess
import sys

import paddle_b200  # noqa: F401
from dist_utils import ROOT


def test_all_alias_paths_import():
    from paddle_b200 import _compat_paths as cp

    assert len(cp._ALIASES) > 100
    for rel in sorted(cp._ALIASES):
        importlib.import_module("paddle_b200." + rel)


def test_common_reference_imports():
    from paddle_b200.autograd.py_layer import PyLayer
    from paddle_b200.distributed.fleet.base.topology import CommunicateTopology, HybridCommunicateGroup  # noqa: F401
    from paddle_b200.distributed.fleet.layers.mpu import mp_ops
    from paddle_b200.distributed.fleet.meta_parallel import ColumnParallelLinear, LayerDesc, PipelineLayer, get_rng_state_tracker  # noqa: F401
    from paddle_b200.distributed.fleet.meta_parallel.sharding.group_sharded_stage3 import GroupShardedStage3  # noqa: F401
    from paddle_b200.distributed.fleet.utils.sequence_parallel_utils import ScatterOp  # noqa: F401
    from paddle_b200.hapi.callbacks import EarlyStopping  # noqa: F401
    from paddle_b200.incubate.distributed.models.moe import MoELayer  # noqa: F401
    from paddle_b200.incubate.distributed.models.moe.gate import GShardGate  # noqa: F401
    from paddle_b200.io.dataloader.batch_sampler import DistributedBatchSampler  # noqa: F401
    from paddle_b200.nn.functional.flash_attention import flash_attention, scaled_dot_product_attention  # noqa: F401
    from paddle_b200.nn.layer.transformer import TransformerEncoderLayer  # noqa: F401
    from paddle_b200.static.amp import AutoMixedPrecisionLists
    from paddle_b200.tensor.math import add
    from paddle_b200.text.datasets import Imdb  # noqa: F401

    assert PyLayer is paddle_b200.autograd.PyLayer and add is paddle_b200.add and hasattr(mp_ops, "_c_identity")
    assert paddle_b200.distributed.fleet.meta_parallel.PipelineLayer is PipelineLayer
    lists = AutoMixedPrecisionLists(custom_white_list=["matmul"], custom_black_list=["softmax"])
    assert "matmul" in lists.white_list


def test_install_as_paddle_single_module_objects():
    code = """
import paddle_b200
paddle_b200.install_as_paddle()
import paddle
import paddle.static as S1
import paddle_b200.static as S2
import paddle.nn.functional as F
from paddle.distributed.fleet.meta_parallel import PipelineLayer
from paddle.vision.models import resnet18
import paddle.distributed.fleet.utils.sequence_parallel_utils as spu
import paddle.optimizer.lr
assert S1 is S2 and paddle.nn is paddle_b200.nn and F is paddle_b200.nn.functional
assert paddle.optimizer.lr.StepDecay is paddle_b200.optimizer.lr.StepDecay
x = paddle.to_tensor([1.0, 2.0])
assert float(paddle.sum(F.relu(x))) == 3.0
print("OK")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]
