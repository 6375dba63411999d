"""Parameter-server mode end to end: 1 server + 2 trainers as separate processes (TRAINING_ROLE env protocol), sparse embedding and a
dense bias living on the server, asynchronous and synchronous pushes, table save / reload.
Parity: test/ps/ + test/legacy_test/test_dist_fleet_*.py (CTR model on the PS)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dist_utils import ROOT, free_port


def _launch(tmp, a_sync, load=False):
    port = free_port()
    base = dict(os.environ, PYTHONPATH=ROOT, PADDLE_PSERVERS_IP_PORT_LIST=f"127.0.0.1:{port}", PADDLE_TRAINERS_NUM="2", OMP_NUM_THREADS="1",
                PS_ASYNC="1" if a_sync else "0", PS_LOAD="1" if load else "0")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_ps_worker.py"), str(tmp)],
                              env=dict(base, TRAINING_ROLE="PSERVER", POD_IP="127.0.0.1", PADDLE_PORT=str(port)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)]
    for i in range(2):
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_ps_worker.py"), str(tmp)],
                                      env=dict(base, TRAINING_ROLE="TRAINER", PADDLE_TRAINER_ID=str(i)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
            o += "\n[TIMEOUT]"
        outs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n------\n".join(o[-2500:] for o in outs)
    return outs


@pytest.mark.parametrize("a_sync", [False, True])
def test_ps_job(tmp_path, a_sync):
    outs = _launch(tmp_path, a_sync)
    assert "server done" in outs[0] and (tmp_path / "worker0.ok").exists() and (tmp_path / "worker1.ok").exists()
    data = np.load(tmp_path / "ckpt" / "ps_tables_0.npz")
    assert data["sparse::ctr_emb::rows"].shape[1] == 8 and 40 <= len(data["sparse::ctr_emb::ids"]) <= 50 and "dense::bias" in data.files
    if not a_sync:
        # restart from the saved tables: the first losses of the second job start where the first one ended
        first_run = float((tmp_path / "worker0.ok").read_text().split()[1])
        _launch(tmp_path, a_sync, load=True)
        resumed_first = float((tmp_path / "worker0.ok").read_text().split()[0])
        assert resumed_first < first_run * 1.6 + 0.1


def test_launch_ps_controller(tmp_path):
    """python -m paddle_b200.distributed.launch --server_num 1 --trainer_num 2 <script>: role env wiring + logs."""
    r = subprocess.run([sys.executable, "-m", "paddle_b200.distributed.launch", "--server_num", "1", "--trainer_num", "2", "--log_dir", str(tmp_path / "log"),
                        os.path.join(ROOT, "tests", "_ps_worker.py"), str(tmp_path)], env=dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="1"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "server done" in (tmp_path / "log" / "serverlog.0").read_text()
    assert (tmp_path / "worker0.ok").exists() and (tmp_path / "worker1.ok").exists()
