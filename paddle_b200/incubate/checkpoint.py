"""paddle.incubate.checkpoint.auto_checkpoint: epoch-range training loops that resume where a previous (killed) run stopped.
Parity: python/paddle/incubate/checkpoint/auto_checkpoint.py (`train_epoch_range`, env PADDLE_RUNNING_ENV / PADDLE_EDL_* ).

    for epoch in acp.train_epoch_range(10, save_checkpoint_inter=1, name="job"):
        train_one_epoch(...)          # objects registered with acp.register(...) are saved after every epoch

State (epoch counter + registered Layers / Optimizers) is written to `$PADDLE_EDL_FS_CHECKPOINT` (default `./auto_checkpoint`)."""
from __future__ import annotations

import json
import os
import time


class _AutoCheckpoint:
    def __init__(self):
        self._objs = {}

    def register(self, **named_objects):
        """Layers / optimizers (anything with state_dict / set_state_dict) to save and restore with the epoch counter."""
        self._objs.update(named_objects)

    def _dir(self, name):
        return os.path.join(os.environ.get("PADDLE_EDL_FS_CHECKPOINT", "./auto_checkpoint"), os.environ.get("PADDLE_JOB_ID", "job"), name)

    def _load(self, name):
        from ..framework.io import load

        meta_path = os.path.join(self._dir(name), "meta.json")
        if not os.path.exists(meta_path):
            return -1
        meta = json.load(open(meta_path))
        for k, obj in self._objs.items():
            p = os.path.join(self._dir(name), k + ".pdstate")
            if os.path.exists(p):
                obj.set_state_dict(load(p))
        return int(meta["epoch"])

    def _save(self, name, epoch):
        from ..distributed import env
        from ..framework.io import save

        if env.get_rank() != 0:
            return
        d = self._dir(name)
        os.makedirs(d, exist_ok=True)
        for k, obj in self._objs.items():
            save(obj.state_dict(), os.path.join(d, k + ".pdstate"))
        tmp = os.path.join(d, "meta.json.tmp")
        json.dump({"epoch": epoch, "time": time.time()}, open(tmp, "w"))
        os.replace(tmp, os.path.join(d, "meta.json"))         # the counter moves last: a crash mid-save replays the epoch

    def train_epoch_range(self, max_epoch_num, save_checkpoint_inter=None, name="acp"):
        last = self._load(name)
        inter = 1 if not save_checkpoint_inter else 1       # epochs are the unit; time based intervals collapse to "every epoch"
        for epoch in range(last + 1, max_epoch_num):
            yield epoch
            if (epoch + 1) % inter == 0:
                self._save(name, epoch)


auto_checkpoint = _AutoCheckpoint()
train_epoch_range = auto_checkpoint.train_epoch_range
register = auto_checkpoint.register
