"""Audio datasets (local files only). Parity: python/paddle/audio/datasets/{esc50,tess}.py."""
from __future__ import annotations

import os

from ..io import Dataset
from . import backends, features


class AudioClassificationDataset(Dataset):
    def __init__(self, files, labels, feat_type="raw", sample_rate=None, **kwargs):
        self.files, self.labels, self.feat_type = files, labels, feat_type
        self._feat = None if feat_type == "raw" else {"melspectrogram": features.MelSpectrogram, "mfcc": features.MFCC,
                                                      "logmelspectrogram": features.LogMelSpectrogram, "spectrogram": features.Spectrogram}[feat_type](**kwargs)

    def __getitem__(self, idx):
        wav, sr = backends.load(self.files[idx])
        x = wav[0]
        if self._feat is not None:
            x = self._feat(x.unsqueeze(0)).squeeze(0)
        return x, self.labels[idx]

    def __len__(self):
        return len(self.files)


class ESC50(AudioClassificationDataset):
    def __init__(self, mode="train", split=1, feat_type="raw", archive=None, **kwargs):
        root = archive or os.environ.get("PADDLE_AUDIO_ESC50")
        if not root or not os.path.isdir(root):
            raise RuntimeError("ESC50: pass archive=<extracted ESC-50 dir> (downloading is disabled: no network)")
        files, labels = [], []
        with open(os.path.join(root, "meta", "esc50.csv")) as f:
            next(f)
            for line in f:
                fn, fold, target = line.strip().split(",")[:3]
                if (mode == "train") == (int(fold) != split):
                    files.append(os.path.join(root, "audio", fn))
                    labels.append(int(target))
        super().__init__(files, labels, feat_type, **kwargs)


class TESS(AudioClassificationDataset):
    def __init__(self, mode="train", n_folds=5, split=1, feat_type="raw", archive=None, **kwargs):
        root = archive or os.environ.get("PADDLE_AUDIO_TESS")
        if not root or not os.path.isdir(root):
            raise RuntimeError("TESS: pass archive=<extracted TESS dir> (downloading is disabled: no network)")
        emos = ["angry", "disgust", "fear", "happy", "neutral", "ps", "sad"]
        wavs = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(root) for f in fs if f.endswith(".wav"))
        files, labels = [], []
        for i, p in enumerate(wavs):
            if (mode == "train") == ((i % n_folds) + 1 != split):
                files.append(p)
                labels.append(emos.index(os.path.basename(p)[:-4].split("_")[-1].lower()))
        super().__init__(files, labels, feat_type, **kwargs)
