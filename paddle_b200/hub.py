"""paddle.hub. Parity: python/paddle/hapi/hub.py (list / help / load from a local directory; github needs network)."""
import importlib.util
import os
import sys

MODULE_HUBCONF = "hubconf.py"


def _import(repo_dir):
    path = os.path.join(repo_dir, MODULE_HUBCONF)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found")
    sys.path.insert(0, repo_dir)
    try:
        spec = importlib.util.spec_from_file_location("hubconf", path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        sys.path.remove(repo_dir)
    return m


def _resolve(repo_dir, source):
    if source not in ("github", "gitee", "local"):
        raise ValueError(f'Unknown source: "{source}". Allowed values: "github" | "gitee" | "local".')
    if source != "local":
        raise RuntimeError("paddle.hub: remote sources need network access; clone the repo and use source='local'")
    return repo_dir


def list(repo_dir, source="github", force_reload=False):  # noqa: A001
    m = _import(_resolve(repo_dir, source))
    return [k for k, v in vars(m).items() if callable(v) and not k.startswith("_")]


def help(repo_dir, model, source="github", force_reload=False):  # noqa: A001
    return getattr(_import(_resolve(repo_dir, source)), model).__doc__


def load(repo_dir, model, source="github", force_reload=False, **kwargs):
    return getattr(_import(_resolve(repo_dir, source)), model)(**kwargs)
