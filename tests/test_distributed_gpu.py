"""Multi-GPU tests (NCCL + peer-memory kernels). Need >= 2 visible GPUs; skipped otherwise."""
import pytest
import torch

from dist_utils import run_dist

pytestmark = pytest.mark.gpu


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def test_p2p_collectives_and_fused_linear():
    _need(2)
    run_dist("p2p_kernels", 2, extra_env={"B200_TEST_GPU": "1"})


def test_mp_sp_parity_gpu():
    _need(2)
    run_dist("mp_sp_parity", 2, extra_env={"B200_TEST_GPU": "1"})


def test_mp_sp_bf16_fused_paths():
    _need(2)
    run_dist("mp_sp_bf16", 2, extra_env={"B200_TEST_GPU": "1"})


def test_pp_gpu():
    _need(2)
    run_dist("pp", 2, extra_env={"B200_TEST_GPU": "1"})


@pytest.mark.parametrize("mode", ["ZBH1", "1F1B"])
def test_pp_bf16_mailbox_schedules(mode):
    """pp=2 bf16 through the peer-memory mailbox; zero-bubble B/W split with arena-fused weight gradients."""
    _need(2)
    run_dist("pp_bf16_zero_bubble", 2, extra_env={"B200_TEST_GPU": "1", "B200_TEST_PP_MODE": mode})


def test_pp_interleave_gpu():
    _need(2)
    run_dist("pp_interleave", 2, extra_env={"B200_TEST_GPU": "1"})


def test_moe_fused_dispatch_combine():
    _need(2)
    run_dist("moe_fused_a2a", 2, extra_env={"B200_TEST_GPU": "1"})


def test_data_parallel_gpu():
    """DataParallel with the gradient slab in the symmetric heap (in-place peer-memory all-reduce per bucket)."""
    _need(2)
    run_dist("dp", 2, extra_env={"B200_TEST_GPU": "1"})


def test_group_sharded_gpu():
    """GroupSharded stage 2/3 over the windowed peer-memory reduce-scatter / all-gather."""
    _need(2)
    run_dist("sharding", 2, extra_env={"B200_TEST_GPU": "1"})


# ---- wider worlds (4 / 8 GPUs): the peer-memory kernels and the fused paths beyond a pair of ranks -------------------------------------
@pytest.mark.parametrize("world", [4, 8])
def test_p2p_collectives_and_fused_linear_wide(world):
    _need(world)
    run_dist("p2p_kernels", world, extra_env={"B200_TEST_GPU": "1"}, timeout=400)


@pytest.mark.parametrize("world", [4, 8])
def test_moe_fused_dispatch_combine_wide(world):
    _need(world)
    run_dist("moe_fused_a2a", world, extra_env={"B200_TEST_GPU": "1"}, timeout=400)


@pytest.mark.parametrize("world", [4, 8])
def test_data_parallel_and_sharding_wide(world):
    _need(world)
    run_dist("dp", world, extra_env={"B200_TEST_GPU": "1"}, timeout=400)
    run_dist("sharding", world, extra_env={"B200_TEST_GPU": "1"}, timeout=400)


def test_hybrid_mp_pp_gpu():
    _need(4)
    run_dist("hybrid_mp_pp", 4, extra_env={"B200_TEST_GPU": "1"}, timeout=400)


def test_stress_dp_allreduce_overlaps_mp_fused_gemms():
    """1000 iterations of a dp-group spinning all-reduce on a side stream under mp-group fused GEMMs (the co-residency / deadlock
    scenario): must neither hang (bounded spins trap after 10 s) nor corrupt results."""
    _need(4)
    run_dist("stress_dp_mp_overlap", 4, extra_env={"B200_TEST_GPU": "1"}, timeout=600)


@pytest.mark.xfail(strict=False, reason="first hardware run of the NVLS (multimem) kernels")
def test_nvls_collectives():
    """NVSwitch multicast all-reduce / reduce-scatter / all-gather (csrc/comm/nvls_collectives.cu) vs NCCL; skipped without NVLS support."""
    _need(2)
    run_dist("nvls_kernels", 2, extra_env={"B200_TEST_GPU": "1"})
