"""paddle.utils.download. Parity: python/paddle/utils/download.py (offline: only local/cached files resolve)."""
import os

WEIGHTS_HOME = os.path.expanduser("~/.cache/paddle/hapi/weights")


def get_weights_path_from_url(url, md5sum=None):
    path = os.path.join(WEIGHTS_HOME, os.path.basename(url))
    if os.path.exists(path):
        return path
    raise RuntimeError(f"cannot download {url}: no network access; place the file at {path}")


def get_path_from_url(url, root_dir, md5sum=None, check_exist=True, decompress=True, method="get"):
    path = os.path.join(root_dir, os.path.basename(url))
    if os.path.exists(path):
        return path
    raise RuntimeError(f"cannot download {url}: no network access; place the file at {path}")
