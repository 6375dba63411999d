"""paddle.dataset (legacy reader-creator datasets). Parity: python/paddle/dataset/{mnist,cifar,uci_housing,imdb,common}.py.
Each `train()/test()` returns a reader creator (a zero-argument callable yielding samples) built on the map-style datasets
of paddle_b200.vision / paddle_b200.text; files are looked up in the local cache (no network in this environment)."""
from . import common  # noqa: F401


def _reader_from(ds_ctor, **kw):
    def reader():
        ds = ds_ctor(**kw)
        for i in range(len(ds)):
            yield ds[i]
    return reader


class _Mod:
    def __init__(self, ctor, train_kw, test_kw):
        self._c, self._tr, self._te = ctor, train_kw, test_kw

    def train(self, **kw):
        return _reader_from(self._c(), **{**self._tr, **kw})

    def test(self, **kw):
        return _reader_from(self._c(), **{**self._te, **kw})


def _lazy(path, name):
    def get():
        import importlib

        return getattr(importlib.import_module(path), name)
    return get


mnist = _Mod(_lazy("paddle_b200.vision.datasets", "MNIST"), {"mode": "train"}, {"mode": "test"})
cifar = _Mod(_lazy("paddle_b200.vision.datasets", "Cifar10"), {"mode": "train"}, {"mode": "test"})
flowers = _Mod(_lazy("paddle_b200.vision.datasets", "Flowers"), {"mode": "train"}, {"mode": "test"})
uci_housing = _Mod(_lazy("paddle_b200.text", "UCIHousing"), {"mode": "train"}, {"mode": "test"})
imdb = _Mod(_lazy("paddle_b200.text", "Imdb"), {"mode": "train"}, {"mode": "test"})
imikolov = _Mod(_lazy("paddle_b200.text", "Imikolov"), {"mode": "train"}, {"mode": "test"})
movielens = _Mod(_lazy("paddle_b200.text", "Movielens"), {"mode": "train"}, {"mode": "test"})
wmt14 = _Mod(_lazy("paddle_b200.text", "WMT14"), {"mode": "train"}, {"mode": "test"})
wmt16 = _Mod(_lazy("paddle_b200.text", "WMT16"), {"mode": "train"}, {"mode": "test"})
conll05 = _Mod(_lazy("paddle_b200.text", "Conll05st"), {}, {})
