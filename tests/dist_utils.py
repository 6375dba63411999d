"""Launch helper for multi-process tests (gloo on CPU, NCCL on GPU). Parity: test/legacy_test/test_dist_base.py."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_dist(case, world, timeout=240, extra_env=None):
    """Runs tests/_dist_worker.py <case> on `world` ranks; raises with the ranks' output on failure."""
    port = free_port()
    procs = []
    for r in range(world):
        e = dict(os.environ)
        e.update({"RANK": str(r), "WORLD_SIZE": str(world), "LOCAL_RANK": str(r), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port),
                  "PYTHONPATH": ROOT + os.pathsep + e.get("PYTHONPATH", ""), "OMP_NUM_THREADS": "1"})
        if extra_env:
            e.update(extra_env)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), case], env=e,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    failed = False
    for p in procs:
        try:
            o, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
            o += "\n[TIMEOUT]"
            failed = True
        outs.append(o)
        failed = failed or p.returncode != 0
    if procs and all(p.returncode == 77 for p in procs):          # the case reported that the machine lacks the feature
        import pytest

        pytest.skip(f"{case}: " + (outs[0].strip().splitlines() or ["unsupported here"])[-1])
    if failed:
        raise AssertionError("distributed case %s failed:\n%s" % (case, "\n------\n".join(o[-3000:] for o in outs)))
    return outs
