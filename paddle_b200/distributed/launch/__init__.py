"""python -m paddle_b200.distributed.launch — collective launcher. Parity: python/paddle/distributed/launch/ (main.py,
controllers/collective.py, job/pod/container, watcher): one process per GPU, per-rank log files, failure watch with
whole-pod teardown, optional elastic restarts (--max_restart)."""
from .main import launch, main  # noqa: F401
