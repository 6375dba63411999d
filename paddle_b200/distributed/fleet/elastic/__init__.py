"""Elastic training manager. Parity: python/paddle/distributed/fleet/elastic/{__init__,manager,collective}.py.

The reference keeps membership in etcd; there is no etcd here, membership lives in the job's TCPStore (the same one the
launcher and the comm watchdog use): every node writes a heartbeat key `elastic/<job>/nodes/<host>` with a timestamp, the
manager on each node reads the live set and decides HOLD (fewer than `np_min` nodes), RESTART (membership changed inside
[np_min, np_max]) or EXIT (more than `np_max` / fatal)."""
from __future__ import annotations

import os
import signal
import subprocess
import threading
import time

ELASTIC_EXIT_CODE = 101
ELASTIC_AUTO_PARALLEL_EXIT_CODE = 102
ELASTIC_TIMEOUT = 2 * 60


class ElasticStatus:
    COMPLETED = "completed"
    ERROR = "error"
    HOLD = "hold"
    RESTART = "restart"
    EXIT = "exit"


class ElasticLevel:
    FAULT_TOLERANCE = 1
    ELASTIC = 2


class _DictStore:
    """In-process stand-in for a TCPStore (tests / single node)."""

    def __init__(self):
        self._d, self._lock = {}, threading.Lock()

    def set(self, k, v):
        with self._lock:
            self._d[k] = v if isinstance(v, bytes) else str(v).encode()

    def get(self, k):
        with self._lock:
            return self._d[k]

    def delete_key(self, k):
        with self._lock:
            return self._d.pop(k, None) is not None

    def keys(self):
        with self._lock:
            return list(self._d)


class LauncherInterface:
    """Owns the local trainer processes. Parity: elastic/manager.py:LauncherInterface."""

    def __init__(self, args):
        self.args, self.procs = args, []

    def launch(self):
        raise NotImplementedError

    def _terminate_procs(self):
        for p in self.procs:
            if p.poll() is None:
                p.send_signal(signal.SIGTERM)
        deadline = time.time() + 10
        for p in self.procs:
            while p.poll() is None and time.time() < deadline:
                time.sleep(0.1)
            if p.poll() is None:
                p.kill()
        self.procs = []

    def _check_procs(self):
        codes = [p.poll() for p in self.procs]
        if any(c is None for c in codes):
            return None
        bad = [c for c in codes if c]
        return bad[0] if bad else 0

    def watch(self):
        return self._check_procs()

    def stop(self):
        self._terminate_procs()


class CollectiveLauncher(LauncherInterface):
    """Starts `args.training_script` once per local device with the collective env. Parity: elastic/collective.py."""

    def launch(self):
        import sys

        nproc = int(getattr(self.args, "nproc_per_node", 1) or 1)
        master = getattr(self.args, "master", None) or "127.0.0.1:29531"
        addr, port = master.rsplit(":", 1)
        for r in range(nproc):
            env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(nproc), MASTER_ADDR=addr, MASTER_PORT=port,
                       PADDLE_TRAINER_ID=str(r), PADDLE_TRAINERS_NUM=str(nproc))
            self.procs.append(subprocess.Popen([sys.executable, self.args.training_script, *getattr(self.args, "training_script_args", [])], env=env))


class ElasticManager:
    def __init__(self, args=None, store=None, job_id=None, np=None, host=None, heartbeat_s=1.0, ttl_s=6.0):
        self.args = args
        self.store = store if store is not None else _DictStore()
        self.job_id = job_id or os.environ.get("PADDLE_ELASTIC_JOB_ID", "default")
        np = np if np is not None else os.environ.get("PADDLE_ELASTIC_NP", "1")
        lo, _, hi = str(np).partition(":")
        self.np_min, self.np_max = int(lo), int(hi or lo)
        self.elastic_level = ElasticLevel.ELASTIC if self.np_max > self.np_min else ElasticLevel.FAULT_TOLERANCE
        self.host = host or os.environ.get("POD_IP", f"127.0.0.1-{os.getpid()}")
        self.heartbeat_s, self.ttl_s = heartbeat_s, ttl_s
        self.enable = self.np_min > 0
        self._prefix = f"elastic/{self.job_id}/nodes/"
        self._stop = threading.Event()
        self._thread = None
        self._last_members = None
        self.launcher = None
        self.stopped = False

    # ---- membership ----
    def _beat(self):
        self.store.set(self._prefix + self.host, repr(time.time()))
        reg = set(self._registry())
        if self.host not in reg:
            reg.add(self.host)
            self.store.set(f"elastic/{self.job_id}/registry", ",".join(sorted(reg)))

    def _registry(self):
        try:
            raw = self.store.get(f"elastic/{self.job_id}/registry")
        except Exception:
            return []
        return [h for h in raw.decode().split(",") if h]

    def hosts(self):
        """Live members (heartbeat younger than ttl)."""
        now, live = time.time(), []
        for h in self._registry():
            try:
                ts = float(self.store.get(self._prefix + h).decode())
            except Exception:
                continue
            if now - ts <= self.ttl_s:
                live.append(h)
        return sorted(live)

    def start_heartbeat(self):
        self._beat()

        def loop():
            while not self._stop.wait(self.heartbeat_s):
                self._beat()
        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()
        return self

    # ---- decisions ----
    def _match(self, members=None):
        members = self.hosts() if members is None else members
        return self.np_min <= len(members) <= self.np_max

    def wait(self, timeout=ELASTIC_TIMEOUT):
        """Block until enough nodes joined. Returns True when the job may start."""
        end = time.time() + timeout
        while time.time() < end:
            if self._match():
                self._last_members = self.hosts()
                return True
            time.sleep(min(self.heartbeat_s, 0.5))
        return False

    def status(self):
        members = self.hosts()
        if len(members) < self.np_min:
            return ElasticStatus.HOLD
        if len(members) > self.np_max:
            return ElasticStatus.EXIT
        if self._last_members is not None and members != self._last_members:
            self._last_members = members
            return ElasticStatus.RESTART
        self._last_members = members
        return ElasticStatus.COMPLETED if self.stopped else "running"

    def run(self, launcher_cls=CollectiveLauncher):
        self.launcher = launcher_cls(self.args)
        self.launcher.launch()

    def watch(self, poll_s=0.5):
        """Supervise: returns an ElasticStatus when the local job ends or membership changes."""
        while True:
            code = self.launcher.watch() if self.launcher is not None else None
            if code is not None:
                self.stopped = True
                if code == 0:
                    return ElasticStatus.COMPLETED
                return ElasticStatus.RESTART if code in (ELASTIC_EXIT_CODE, ELASTIC_AUTO_PARALLEL_EXIT_CODE) or self.elastic_level else ElasticStatus.ERROR
            st = self.status()
            if st in (ElasticStatus.HOLD, ElasticStatus.RESTART, ElasticStatus.EXIT):
                if self.launcher is not None:
                    self.launcher.stop()
                return st
            time.sleep(poll_s)

    def exit(self, completed=False):
        self._stop.set()
        if self.launcher is not None:
            self.launcher.stop()
        self.store.delete_key(self._prefix + self.host) if hasattr(self.store, "delete_key") else None
        reg = [h for h in self._registry() if h != self.host]
        self.store.set(f"elastic/{self.job_id}/registry", ",".join(reg))


def enable_elastic(args, distribute_mode=None):
    return bool(os.environ.get("PADDLE_ELASTIC_NP") or getattr(args, "elastic_level", 0) or ":" in str(getattr(args, "nnodes", "")))


def launch_elastic(args, distribute_mode=None, store=None):
    """Run-until-complete loop with restarts on membership change. Parity: elastic/__init__.py:launch_elastic."""
    mgr = ElasticManager(args, store=store, np=getattr(args, "nnodes", None) or os.environ.get("PADDLE_ELASTIC_NP", "1")).start_heartbeat()
    try:
        while True:
            if not mgr.wait():
                return ElasticStatus.ERROR
            mgr.run()
            st = mgr.watch()
            if st in (ElasticStatus.COMPLETED, ElasticStatus.ERROR, ElasticStatus.EXIT):
                return st
    finally:
        mgr.exit()
