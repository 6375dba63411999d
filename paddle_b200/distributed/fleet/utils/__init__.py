"""paddle.distributed.fleet.utils. Parity: python/paddle/distributed/fleet/utils/__init__.py (recompute, LocalFS, HDFSClient,
DistributedInfer) and its sub-modules (sequence_parallel_utils, hybrid_parallel_util, mix_precision_utils, tensor_fusion_helper)."""
from __future__ import annotations

import os
import shutil

from ..recompute import recompute, recompute_hybrid, recompute_sequential  # noqa: F401
from . import hybrid_parallel_util, mix_precision_utils, sequence_parallel_utils, tensor_fusion_helper  # noqa: F401


class LocalFS:
    """Local file system client. Parity: fleet/utils/fs.py:LocalFS."""

    def ls_dir(self, fs_path):
        if not self.is_exist(fs_path):
            return [], []
        dirs, files = [], []
        for f in os.listdir(fs_path):
            (dirs if os.path.isdir(os.path.join(fs_path, f)) else files).append(f)
        return dirs, files

    def mkdirs(self, fs_path):
        os.makedirs(fs_path, exist_ok=True)

    def rename(self, src, dst):
        os.rename(src, dst)

    def delete(self, fs_path):
        if os.path.isdir(fs_path):
            shutil.rmtree(fs_path)
        elif os.path.exists(fs_path):
            os.remove(fs_path)

    def is_file(self, fs_path):
        return os.path.isfile(fs_path)

    def is_dir(self, fs_path):
        return os.path.isdir(fs_path)

    def is_exist(self, fs_path):
        return os.path.exists(fs_path)

    def touch(self, fs_path, exist_ok=True):
        if os.path.exists(fs_path) and not exist_ok:
            raise FileExistsError(fs_path)
        open(fs_path, "a").close()

    def mv(self, src, dst, overwrite=False, test_exists=False):
        if overwrite and os.path.exists(dst):
            self.delete(dst)
        shutil.move(src, dst)

    def list_dirs(self, fs_path):
        return self.ls_dir(fs_path)[0]

    def need_upload_download(self):
        return False


class HDFSClient(LocalFS):
    """HDFS client shell wrapper. There is no hadoop binary (nor network) in this environment: construction succeeds so
    configuration code can run, every remote operation raises."""

    def __init__(self, hadoop_home=None, configs=None, time_out=300000, sleep_inter=1000):
        self._home, self._configs = hadoop_home, configs or {}

    def _unavailable(self, *a, **k):
        raise RuntimeError("HDFSClient: no hadoop installation is available in this environment")

    ls_dir = mkdirs = rename = delete = touch = mv = upload = download = _unavailable

    def need_upload_download(self):
        return True


class DistributedInfer:
    """Parity: fleet/utils/ps_util.py:DistributedInfer — PS-mode inference helper; with in-process tables it just runs the program."""

    def __init__(self, main_program=None, startup_program=None):
        self.main_program, self.startup_program = main_program, startup_program

    def init_distributed_infer_env(self, exe, loss, role_maker=None, dirname=None):
        if self.startup_program is not None:
            exe.run(self.startup_program)

    def get_dist_infer_program(self):
        return self.main_program
