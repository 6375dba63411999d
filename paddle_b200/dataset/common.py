"""paddle.dataset.common. Parity: python/paddle/dataset/common.py."""
import glob
import hashlib
import os
import pickle

DATA_HOME = os.path.expanduser(os.environ.get("PADDLE_DATA_HOME", "~/.cache/paddle/dataset"))


def md5file(fname):
    h = hashlib.md5()
    with open(fname, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def download(url, module_name, md5sum, save_name=None):
    d = os.path.join(DATA_HOME, module_name)
    os.makedirs(d, exist_ok=True)
    fn = os.path.join(d, save_name or url.split("/")[-1])
    if os.path.exists(fn) and (md5sum is None or md5file(fn) == md5sum):
        return fn
    from ..utils.download import get_path_from_url

    return get_path_from_url(url, d, md5sum)


def split(reader, line_count, suffix="%05d.pickle", dumper=pickle.dump):
    lines, idx = [], 0
    for i, d in enumerate(reader()):
        lines.append(d)
        if (i + 1) % line_count == 0:
            with open(suffix % idx, "wb") as f:
                dumper(lines, f)
            lines, idx = [], idx + 1
    if lines:
        with open(suffix % idx, "wb") as f:
            dumper(lines, f)


def cluster_files_reader(files_pattern, trainer_count, trainer_id, loader=pickle.load):
    def reader():
        files = sorted(glob.glob(files_pattern))
        for i, fn in enumerate(files):
            if i % trainer_count == trainer_id:
                with open(fn, "rb") as f:
                    yield from loader(f)
    return reader
