"""paddle.vision.datasets. Parity: python/paddle/vision/datasets/{mnist,cifar,flowers,voc2012,folder}.py.
No network here: datasets read local files (same on-disk formats as the reference) and raise a clear error otherwise."""
from __future__ import annotations

import gzip
import os
import pickle
import struct
import tarfile

import numpy as np

from ...io import Dataset

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")


def _need(path, what):
    if path is None or not os.path.exists(path):
        raise RuntimeError(f"{what}: file '{path}' not found and downloading is disabled (no network); pass a local path")


def _pil_loader(path):
    from PIL import Image

    with open(path, "rb") as f:
        return Image.open(f).convert("RGB")


def default_loader(path):
    return _pil_loader(path)


class MNIST(Dataset):
    NAME = "mnist"

    def __init__(self, image_path=None, label_path=None, mode="train", transform=None, download=True, backend=None):
        assert mode.lower() in ("train", "test")
        self.mode, self.transform, self.backend = mode.lower(), transform, backend or "pil"
        _need(image_path, type(self).__name__)
        _need(label_path, type(self).__name__)
        with gzip.open(image_path, "rb") as f:
            magic, n, rows, cols = struct.unpack(">IIII", f.read(16))
            self.images = np.frombuffer(f.read(), dtype=np.uint8).reshape(n, rows, cols)
        with gzip.open(label_path, "rb") as f:
            magic, n = struct.unpack(">II", f.read(8))
            self.labels = np.frombuffer(f.read(), dtype=np.uint8).astype(np.int64)

    def __getitem__(self, idx):
        img, label = self.images[idx], np.array([self.labels[idx]]).astype("int64")
        if self.backend == "pil":
            from PIL import Image

            img = Image.fromarray(img, mode="L")
        if self.transform is not None:
            img = self.transform(img)
        return (img, label) if self.backend == "pil" else (np.asarray(img, dtype="float32"), label)

    def __len__(self):
        return len(self.labels)


class FashionMNIST(MNIST):
    NAME = "fashion-mnist"


class Cifar10(Dataset):
    _train = ["data_batch"]
    _test = ["test_batch"]
    _label_key = b"labels"

    def __init__(self, data_file=None, mode="train", transform=None, download=True, backend=None):
        assert mode.lower() in ("train", "test")
        self.mode, self.transform, self.backend = mode.lower(), transform, backend or "pil"
        _need(data_file, type(self).__name__)
        names = self._train if self.mode == "train" else self._test
        self.data = []
        with tarfile.open(data_file, mode="r") as f:
            for member in f.getmembers():
                if any(n in member.name for n in names):
                    batch = pickle.load(f.extractfile(member), encoding="bytes")
                    labels = batch.get(self._label_key, batch.get(b"fine_labels"))
                    for x, y in zip(batch[b"data"], labels):
                        self.data.append((x, int(y)))

    def __getitem__(self, idx):
        x, y = self.data[idx]
        img = np.reshape(x, [3, 32, 32]).transpose(1, 2, 0)
        if self.backend == "pil":
            from PIL import Image

            img = Image.fromarray(img.astype("uint8"))
        if self.transform is not None:
            img = self.transform(img)
        return img, np.array(y).astype("int64")

    def __len__(self):
        return len(self.data)


class Cifar100(Cifar10):
    _train = ["train"]
    _test = ["test"]
    _label_key = b"fine_labels"


class DatasetFolder(Dataset):
    def __init__(self, root, loader=None, extensions=None, transform=None, is_valid_file=None):
        self.root, self.transform = root, transform
        self.loader = loader or default_loader
        extensions = extensions or IMG_EXTENSIONS
        classes = sorted(d.name for d in os.scandir(root) if d.is_dir())
        self.classes = classes
        self.class_to_idx = {c: i for i, c in enumerate(classes)}
        valid = is_valid_file or (lambda p: p.lower().endswith(tuple(extensions)))
        self.samples = []
        for c in classes:
            for dp, _, fns in sorted(os.walk(os.path.join(root, c), followlinks=True)):
                for fn in sorted(fns):
                    p = os.path.join(dp, fn)
                    if valid(p):
                        self.samples.append((p, self.class_to_idx[c]))
        if not self.samples:
            raise RuntimeError(f"Found 0 files in subfolders of: {root}")
        self.targets = [s[1] for s in self.samples]

    def __getitem__(self, index):
        path, target = self.samples[index]
        sample = self.loader(path)
        if self.transform is not None:
            sample = self.transform(sample)
        return sample, target

    def __len__(self):
        return len(self.samples)


class ImageFolder(Dataset):
    def __init__(self, root, loader=None, extensions=None, transform=None, is_valid_file=None):
        self.root, self.transform = root, transform
        self.loader = loader or default_loader
        extensions = extensions or IMG_EXTENSIONS
        valid = is_valid_file or (lambda p: p.lower().endswith(tuple(extensions)))
        self.samples = []
        for dp, _, fns in sorted(os.walk(root, followlinks=True)):
            for fn in sorted(fns):
                p = os.path.join(dp, fn)
                if valid(p):
                    self.samples.append(p)
        if not self.samples:
            raise RuntimeError(f"Found 0 files in: {root}")

    def __getitem__(self, index):
        sample = self.loader(self.samples[index])
        if self.transform is not None:
            sample = self.transform(sample)
        return [sample]

    def __len__(self):
        return len(self.samples)


class Flowers(Dataset):
    def __init__(self, data_file=None, label_file=None, setid_file=None, mode="train", transform=None, download=True, backend=None):
        import scipy.io as sio

        for p in (data_file, label_file, setid_file):
            _need(p, "Flowers")
        self.transform, self.backend = transform, backend or "pil"
        self.labels = sio.loadmat(label_file)["labels"][0]
        key = {"train": "tstid", "valid": "valid", "test": "trnid"}[mode.lower()]
        self.indexes = sio.loadmat(setid_file)[key][0]
        self.data_tar = tarfile.open(data_file)
        self.name2mem = {m.name: m for m in self.data_tar.getmembers()}

    def __getitem__(self, idx):
        from PIL import Image

        index = self.indexes[idx]
        img = Image.open(self.data_tar.extractfile(self.name2mem["jpg/image_%05d.jpg" % index])).convert("RGB")
        if self.backend != "pil":
            img = np.asarray(img)
        if self.transform is not None:
            img = self.transform(img)
        return img, np.array([self.labels[index - 1]]).astype("int64")

    def __len__(self):
        return len(self.indexes)


class VOC2012(Dataset):
    def __init__(self, data_file=None, mode="train", transform=None, download=True, backend=None):
        _need(data_file, "VOC2012")
        self.transform, self.backend = transform, backend or "pil"
        self.data_tar = tarfile.open(data_file)
        self.name2mem = {m.name: m for m in self.data_tar.getmembers()}
        flag = {"train": "train", "valid": "val", "test": "val"}[mode.lower()]
        set_file = f"VOCdevkit/VOC2012/ImageSets/Segmentation/{flag}.txt"
        self.ids = [l.strip().decode() for l in self.data_tar.extractfile(self.name2mem[set_file])]

    def __getitem__(self, idx):
        from PIL import Image

        i = self.ids[idx]
        img = Image.open(self.data_tar.extractfile(self.name2mem[f"VOCdevkit/VOC2012/JPEGImages/{i}.jpg"])).convert("RGB")
        lab = Image.open(self.data_tar.extractfile(self.name2mem[f"VOCdevkit/VOC2012/SegmentationClass/{i}.png"]))
        if self.backend != "pil":
            img, lab = np.asarray(img), np.asarray(lab)
        if self.transform is not None:
            img = self.transform(img)
        return img, lab

    def __len__(self):
        return len(self.ids)
