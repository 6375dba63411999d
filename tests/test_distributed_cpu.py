"""Multi-process (gloo, CPU) tests of the distributed stack: collectives, mp layers, sequence parallel, DataParallel,
pipeline 1F1B, hybrid mp x pp, group sharded stage 2/3 — each checks parity with single-process training."""
import pytest

from dist_utils import run_dist


@pytest.mark.parametrize("case,world", [("collectives", 2), ("mp_layers", 2), ("sequence_parallel", 2), ("dp", 2), ("pp", 2),
                                        ("sharding", 2), ("mp_sp_parity", 2), ("hybrid_mp_pp", 4)])
def test_dist_case(case, world):
    run_dist(case, world)
