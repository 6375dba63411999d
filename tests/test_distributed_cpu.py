"""Multi-process (gloo, CPU) tests of the distributed stack: collectives, mp layers, sequence parallel, DataParallel,
pipeline 1F1B, hybrid mp x pp, group sharded stage 2/3 — each checks parity with single-process training."""
import pytest

from dist_utils import run_dist


@pytest.mark.parametrize("case,world", [("collectives", 2), ("mp_layers", 2), ("sequence_parallel", 2), ("dp", 2), ("pp", 2),
                                        ("sharding", 2), ("mp_sp_parity", 2), ("hybrid_mp_pp", 4), ("auto_parallel", 2), ("spmd_rules", 2),
                                        ("dist_checkpoint", 2), ("dist_checkpoint_mp", 2), ("pp_interleave", 2), ("moe_ep", 2), ("dp_no_sync", 2), ("sep_parallel", 2), ("pp_shared_embedding", 2),
                                        ("dp_unused_params", 2), ("hybrid_scaler", 2), ("hybrid_scaler_dp", 2), ("fleet_sharding_degree", 2), ("auto_engine", 2), ("recompute_hybrid_partition", 2), ("hapi_fit", 2)])
def test_dist_case(case, world, tmp_path):
    run_dist(case, world, extra_env={"B200_TEST_TMP": str(tmp_path)})


@pytest.mark.parametrize("case,mode", [("pp", "ZBH1"), ("pp", "FThenB"), ("pp_interleave", "FThenB")])
def test_pipeline_schedules(case, mode, tmp_path):
    """Zero-bubble (B / W split), FThenB and interleaved-FThenB schedules reach the same weights as single-process training."""
    run_dist(case, 2, extra_env={"B200_TEST_TMP": str(tmp_path), "B200_TEST_PP_MODE": mode})


def test_pipeline_schedule_builder():
    """Every schedule is deadlock-free under asynchronous sends; zero-bubble beats 1F1B; the interleaved ramp is 1/V of the plain one."""
    from paddle_b200.distributed.fleet import pp_schedule as P

    for S, M in [(2, 8), (4, 8), (4, 16), (8, 16)]:
        base = P.simulate(P.build("1F1B", S, M), S, merged_w=True)[0]
        zb = P.simulate(P.build("ZBH1", S, M), S)[0]
        assert zb < base and zb >= 3 * M, (S, M, zb, base)
        assert P.simulate(P.build("FThenB", S, M), S, merged_w=True)[0] == base
        for ops in P.build("ZBH1", S, M):
            assert sorted(ops) == sorted([(k, 0, m) for k in "FBW" for m in range(M)])
        v2 = P.simulate(P.build("VPP", S, M, V=2), S, V=2, cost=(0.5, 0.5, 0.5), merged_w=True)[0]
        assert v2 < base
    assert P.simulate(P.build("ZBH1", 2, 8), 2)[0] == 25.0      # 3 M + 1: one forward of ramp is all that is left
    with pytest.raises(ValueError):
        P.build("VPP", 4, 6, V=2)


def test_launch_module(tmp_path):
    """python -m paddle_b200.distributed.launch: per-rank env + logs, failure propagation. Parity: test/collective/test_launch*.sh."""
    import os
    import subprocess
    import sys

    from dist_utils import ROOT

    ok = tmp_path / "ok.py"
    ok.write_text("import os\nprint('R', os.environ['PADDLE_TRAINER_ID'], os.environ['PADDLE_TRAINERS_NUM'], os.environ['MASTER_ADDR'])\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "paddle_b200.distributed.launch", "--nproc_per_node", "2", "--log_dir", str(tmp_path / "log"), str(ok)],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "R 0 2 127.0.0.1" in r.stdout
    assert "R 1 2" in (tmp_path / "log" / "workerlog.1").read_text()
    bad = tmp_path / "bad.py"
    bad.write_text("import os, sys, time\nif os.environ['RANK'] == '1':\n    sys.exit(3)\ntime.sleep(30)\n")
    r = subprocess.run([sys.executable, "-m", "paddle_b200.distributed.launch", "--nproc_per_node", "2", "--log_dir", str(tmp_path / "log2"), str(bad)],
                       env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 3   # the failing rank tears the pod down well before rank 0's sleep ends


def _spawn_fn(path):
    import os

    import paddle_b200.distributed as dist

    dist.init_parallel_env()
    import torch

    t = torch.ones(1) * (dist.get_rank() + 1)
    dist.all_reduce(t)
    open(os.path.join(path, f"r{dist.get_rank()}"), "w").write(str(float(t)))


def test_spawn(tmp_path):
    import paddle_b200.distributed as dist

    dist.spawn(_spawn_fn, args=(str(tmp_path),), nprocs=2)
    assert (tmp_path / "r0").read_text() == "3.0" and (tmp_path / "r1").read_text() == "3.0"


def test_auto_tuner_and_cost_model():
    from paddle_b200.distributed import auto_tuner

    res = auto_tuner.search(num_gpus=8, hidden=5120, layers=40, ffn=13824, vocab=32000, seq=4096, global_batch=32, heads=40)
    assert res and all(c["dp"] * c["mp"] * c["pp"] * c["sharding"] == 8 for c in res)
    assert all(c["mem_gb"] < 180 for c in res) and res[0]["est_ms"] <= res[-1]["est_ms"]
    t = auto_tuner.AutoTuner(dict(num_gpus=8, hidden=5120, layers=40, ffn=13824, vocab=32000, seq=4096, global_batch=32, heads=40))
    a, b = t.search_once(), t.search_once()
    assert a is not None and a != b


def test_parameter_server_tables():
    import numpy as np

    from paddle_b200.distributed import ps

    server = ps.ParameterServer()
    server.create_sparse("emb", 4, optimizer="adagrad", lr=0.1)
    w = ps.Worker(local=server)
    rows = w.pull_sparse("emb", [3, 7, 3])
    assert rows.shape == (3, 4) and np.allclose(rows[0], rows[2])
    w.push_sparse("emb", [3], np.ones((1, 4), np.float32))
    assert not np.allclose(w.pull_sparse("emb", [3])[0], rows[0]) and server.tables["emb"].size() == 2


def test_watchdog_detects_stall():
    import time

    from paddle_b200.distributed.watchdog import CommWatchdog

    hits = []
    wd = CommWatchdog(timeout_s=0.3, interval_s=0.1, on_failure=hits.append).start()
    s = wd.record("all_reduce")
    time.sleep(0.8)
    wd.stop()
    assert hits and "no training progress" in hits[0] and "all_reduce" in hits[0]
    wd.done(s)
