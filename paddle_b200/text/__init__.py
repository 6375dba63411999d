"""paddle.text. Parity: python/paddle/text/{viterbi_decode.py, datasets/*}."""
from __future__ import annotations

import os
import re
import tarfile

import numpy as np
import torch

from ..io import Dataset
from ..nn.layer import Layer
from ..tensor import Tensor


def _raw(t):
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


def viterbi_decode(potentials, transition_params, lengths, include_bos_eos_tag=True, name=None):
    """Parity: python/paddle/text/viterbi_decode.py (viterbi_decode kernel)."""
    pot, trans, lens = _raw(potentials).float(), _raw(transition_params).float(), _raw(lengths).long()
    B, T, N = pot.shape
    if include_bos_eos_tag:
        start, stop = trans[-2], trans[:, -1]   # BOS row = N-2, EOS column = N-1 (reference layout)
        alpha = pot[:, 0] + start[None]
    else:
        alpha = pot[:, 0].clone()
    hist = []
    for t in range(1, T):
        score = alpha[:, :, None] + trans[None]
        best, idx = score.max(1)
        new = best + pot[:, t]
        mask = (t < lens)[:, None]
        alpha = torch.where(mask, new, alpha)
        hist.append(torch.where(mask, idx, torch.arange(N, device=pot.device)[None].expand(B, N)))
    if include_bos_eos_tag:
        alpha = alpha + stop[None]
    scores, last = alpha.max(1)
    path = [last]
    for idx in reversed(hist):
        last = torch.gather(idx, 1, last[:, None]).squeeze(1)
        path.append(last)
    path = torch.stack(path[::-1], 1)
    pos = torch.arange(T, device=pot.device)[None]
    path = torch.where(pos < lens[:, None], path, torch.zeros_like(path))
    return scores.as_subclass(Tensor), path.as_subclass(Tensor)


class ViterbiDecoder(Layer):
    def __init__(self, transitions, include_bos_eos_tag=True, name=None):
        super().__init__()
        self.transitions, self.include_bos_eos_tag = transitions, include_bos_eos_tag

    def forward(self, potentials, lengths):
        return viterbi_decode(potentials, self.transitions, lengths, self.include_bos_eos_tag)


def _need(path, what):
    if path is None or not os.path.exists(path):
        raise RuntimeError(f"{what}: data_file '{path}' not found and downloading is disabled (no network)")


class UCIHousing(Dataset):
    def __init__(self, data_file=None, mode="train", download=True):
        _need(data_file, "UCIHousing")
        data = np.fromfile(data_file, sep=" ").reshape(-1, 14)
        mx, mn, avg = data.max(0), data.min(0), data.mean(0)
        for i in range(13):
            data[:, i] = (data[:, i] - avg[i]) / (mx[i] - mn[i])
        off = int(data.shape[0] * 0.8)
        self.data = (data[:off] if mode == "train" else data[off:]).astype("float32")

    def __getitem__(self, idx):
        return self.data[idx, :-1], self.data[idx, -1:]

    def __len__(self):
        return len(self.data)


class Imdb(Dataset):
    def __init__(self, data_file=None, mode="train", cutoff=150, download=True):
        import re
        import string

        _need(data_file, "Imdb")
        self.docs, self.labels = [], []
        pat = re.compile(rf"aclImdb/{mode}/(pos|neg)/.*\.txt$")
        freq = {}
        texts = []
        with tarfile.open(data_file) as tf:
            for m in tf:
                mt = pat.match(m.name)
                if mt:
                    words = tf.extractfile(m).read().decode("latin-1").lower().translate(str.maketrans("", "", string.punctuation)).split()
                    texts.append((words, 0 if mt.group(1) == "pos" else 1))
                    for w in words:
                        freq[w] = freq.get(w, 0) + 1
        vocab = sorted([w for w, c in freq.items() if c > cutoff], key=lambda w: (-freq[w], w))
        self.word_idx = {w: i for i, w in enumerate(vocab)}
        self.word_idx["<unk>"] = len(vocab)
        unk = self.word_idx["<unk>"]
        for words, lab in texts:
            self.docs.append(np.array([self.word_idx.get(w, unk) for w in words]))
            self.labels.append(np.array([lab]))

    def __getitem__(self, idx):
        return self.docs[idx], self.labels[idx]

    def __len__(self):
        return len(self.docs)


class Imikolov(Dataset):
    def __init__(self, data_file=None, data_type="NGRAM", window_size=-1, mode="train", min_word_freq=50, download=True):
        _need(data_file, "Imikolov")
        name = {"train": "./simple-examples/data/ptb.train.txt", "test": "./simple-examples/data/ptb.valid.txt"}[mode]
        with tarfile.open(data_file) as tf:
            train = tf.extractfile("./simple-examples/data/ptb.train.txt").read().decode().splitlines()
            lines = tf.extractfile(name).read().decode().splitlines()
        freq = {}
        for l in train:
            for w in l.split():
                freq[w] = freq.get(w, 0) + 1
        vocab = sorted([w for w, c in freq.items() if c > min_word_freq], key=lambda w: (-freq[w], w))
        self.word_idx = {w: i for i, w in enumerate(vocab)}
        for tok in ("<unk>", "<s>", "<e>"):
            self.word_idx.setdefault(tok, len(self.word_idx))
        unk = self.word_idx["<unk>"]
        self.data = []
        for l in lines:
            ids = [self.word_idx["<s>"]] + [self.word_idx.get(w, unk) for w in l.split()] + [self.word_idx["<e>"]]
            if data_type == "NGRAM":
                for i in range(window_size, len(ids) + 1):
                    self.data.append(tuple(ids[i - window_size:i]))
            else:
                self.data.append((ids[:-1], ids[1:]))

    def __getitem__(self, idx):
        return tuple(np.array(d) for d in self.data[idx])

    def __len__(self):
        return len(self.data)


class Movielens(Dataset):
    """MovieLens-1M from a local `ml-1m.zip`. Parity: python/paddle/text/datasets/movielens.py.
    Sample: ([uid], [gender], [age_idx], [job], [mid], [category ids], [title word ids], [rating * 2 - 5])."""

    _AGES = [1, 18, 25, 35, 45, 50, 56]

    def __init__(self, data_file=None, mode="train", test_ratio=0.1, rand_seed=0, download=True):
        import zipfile

        _need(data_file, "Movielens")
        self.movie_info, self.user_info = {}, {}
        self.categories_dict, self.movie_title_dict = {}, {}
        title_pat = re.compile(r"^(.*)\((\d+)\)$")
        with zipfile.ZipFile(data_file) as z:
            for line in z.read("ml-1m/movies.dat").decode("latin-1").splitlines():
                if not line.strip():
                    continue
                mid, title, cats = line.strip().split("::")
                m = title_pat.match(title)
                title = (m.group(1) if m else title).strip()
                cats = cats.split("|")
                for c in cats:
                    self.categories_dict.setdefault(c, len(self.categories_dict))
                words = title.lower().split()
                for w in words:
                    self.movie_title_dict.setdefault(w, len(self.movie_title_dict))
                self.movie_info[int(mid)] = (int(mid), [self.categories_dict[c] for c in cats], [self.movie_title_dict[w] for w in words])
            for line in z.read("ml-1m/users.dat").decode("latin-1").splitlines():
                if not line.strip():
                    continue
                uid, gender, age, job = line.strip().split("::")[:4]
                self.user_info[int(uid)] = (int(uid), 0 if gender == "M" else 1, self._AGES.index(int(age)), int(job))
            rnd = np.random.RandomState(rand_seed)
            self.data = []
            for line in z.read("ml-1m/ratings.dat").decode("latin-1").splitlines():
                if not line.strip():
                    continue
                is_test = rnd.rand() < test_ratio
                if is_test != (mode == "test"):
                    continue
                uid, mid, rating = line.strip().split("::")[:3]
                u, mv = self.user_info[int(uid)], self.movie_info[int(mid)]
                self.data.append(([u[0]], [u[1]], [u[2]], [u[3]], [mv[0]], mv[1], mv[2], [float(rating) * 2 - 5.0]))

    def __getitem__(self, idx):
        return tuple(np.array(d) for d in self.data[idx])

    def __len__(self):
        return len(self.data)


class Conll05st(Dataset):
    """CoNLL-2005 SRL test split from a local `conll05st-tests.tar.gz` plus dictionary files.
    Parity: python/paddle/text/datasets/conll05.py. Sample (9 arrays of sentence length): word ids, five predicate-context
    word ids (n2, n1, 0, p1, p2), predicate id, predicate-window mark, BIO label ids."""

    UNK_IDX = 0

    def __init__(self, data_file=None, word_dict_file=None, verb_dict_file=None, target_dict_file=None, emb_file=None, download=True):
        import gzip

        for p, w in ((data_file, "data_file"), (word_dict_file, "word_dict_file"), (verb_dict_file, "verb_dict_file"), (target_dict_file, "target_dict_file")):
            _need(p, f"Conll05st {w}")
        self.word_dict, self.predicate_dict = self._load_dict(word_dict_file), self._load_dict(verb_dict_file)
        self.label_dict = self._load_label_dict(target_dict_file)
        self.emb_file = emb_file
        with tarfile.open(data_file) as tf:
            words = gzip.decompress(tf.extractfile("conll05st-release/test.wsj/words/test.wsj.words.gz").read()).decode().splitlines()
            props = gzip.decompress(tf.extractfile("conll05st-release/test.wsj/props/test.wsj.props.gz").read()).decode().splitlines()
        self.sentences, self.predicates, self.labels = [], [], []
        sent, cols = [], []
        for w, p in zip(words + [""], props + [""]):
            w, p = w.strip(), p.strip().split()
            if w:
                sent.append(w)
                cols.append(p)
                continue
            if sent:
                verbs = [c[0] for c in cols if c[0] != "-"]
                for k, verb in enumerate(verbs):
                    self.sentences.append(sent)
                    self.predicates.append(verb)
                    self.labels.append(self._bio([c[k + 1] for c in cols]))
            sent, cols = [], []

    @staticmethod
    def _bio(col):
        """'(A0*' '*' '*)' bracket column -> B-/I-/O tags."""
        out, cur = [], None
        for tok in col:
            if tok.startswith("("):
                tag = tok[1:tok.index("*")]
                out.append("B-" + tag)
                cur = None if tok.endswith(")") else tag
            elif cur is not None:
                out.append("I-" + cur)
                if tok.endswith(")"):
                    cur = None
            else:
                out.append("O")
        return out

    @staticmethod
    def _load_dict(path):
        with open(path) as f:
            return {l.strip(): i for i, l in enumerate(f) if l.strip()}

    @staticmethod
    def _load_label_dict(path):
        tags = set()
        with open(path) as f:
            for l in f:
                l = l.strip()
                if l.startswith(("B-", "I-")):
                    tags.add(l[2:])
        d = {}
        for t in sorted(tags):
            d["B-" + t] = len(d)
            d["I-" + t] = len(d)
        d["O"] = len(d)
        return d

    def get_dict(self):
        return self.word_dict, self.predicate_dict, self.label_dict

    def get_embedding(self):
        return self.emb_file

    def __getitem__(self, idx):
        sent, verb, labels = self.sentences[idx], self.predicates[idx], self.labels[idx]
        n = len(sent)
        vi = next(i for i, l in enumerate(labels) if l == "B-V") if "B-V" in labels else 0
        mark = [0] * n
        ctx = []
        for d in (-2, -1, 0, 1, 2):
            j = vi + d
            if 0 <= j < n:
                mark[j] = 1
                ctx.append(sent[j])
            else:
                ctx.append("bos" if d < 0 else "eos")
        wid = [self.word_dict.get(w, self.UNK_IDX) for w in sent]
        ctx_ids = [[self.word_dict.get(c, self.UNK_IDX)] * n for c in ctx]
        pred = [self.predicate_dict.get(verb, 0)] * n
        lab = [self.label_dict.get(l, self.label_dict["O"]) for l in labels]
        return tuple(np.array(a) for a in (wid, *ctx_ids, pred, mark, lab))

    def __len__(self):
        return len(self.sentences)



class WMT14(Dataset):
    """WMT14 en-fr subset from a local tar: members `*/src.dict`, `*/trg.dict` and `train/train` | `test/test` | `gen/gen` with
    `source<TAB>target` lines. Parity: python/paddle/text/datasets/wmt14.py. Sample: (src ids, <s>+trg ids, trg ids+<e>)."""

    START, END, UNK, UNK_IDX = "<s>", "<e>", "<unk>", 2

    def __init__(self, data_file=None, mode="train", dict_size=-1, download=True):
        _need(data_file, "WMT14")
        assert mode in ("train", "test", "gen")
        self.src_dict, self.trg_dict, self.data = {}, {}, []
        with tarfile.open(data_file) as tf:
            members = {m.name: m for m in tf if m.isfile()}

            def load_dict(suffix):
                name = next(n for n in members if n.endswith(suffix))
                d = {}
                for i, l in enumerate(tf.extractfile(members[name]).read().decode().splitlines()):
                    if dict_size > 0 and i >= dict_size:
                        break
                    d[l.strip()] = i
                return d

            self.src_dict, self.trg_dict = load_dict("src.dict"), load_dict("trg.dict")
            name = next(n for n in members if n.endswith(f"{mode}/{mode}"))
            for l in tf.extractfile(members[name]).read().decode().splitlines():
                parts = l.rstrip("\n").split("\t")
                if len(parts) != 2:
                    continue
                src = [self.src_dict.get(w, self.UNK_IDX) for w in [self.START, *parts[0].split(), self.END]]
                trg = [self.trg_dict.get(w, self.UNK_IDX) for w in parts[1].split()]
                if len(src) > 80 or len(trg) > 80:
                    continue
                self.data.append((src, [self.trg_dict[self.START], *trg], [*trg, self.trg_dict[self.END]]))

    def get_dict(self, reverse=False):
        if reverse:
            return {v: k for k, v in self.src_dict.items()}, {v: k for k, v in self.trg_dict.items()}
        return self.src_dict, self.trg_dict

    def __getitem__(self, idx):
        return tuple(np.array(d) for d in self.data[idx])

    def __len__(self):
        return len(self.data)


class WMT16(Dataset):
    """WMT16 en-de (Multi30k) from a local tar with `wmt16/train`, `wmt16/val`, `wmt16/test` (`en<TAB>de` lines). Dictionaries are
    built from the training split by frequency, behind `<s>`, `<e>`, `<unk>`. Parity: python/paddle/text/datasets/wmt16.py."""

    START, END, UNK = "<s>", "<e>", "<unk>"

    def __init__(self, data_file=None, mode="train", src_dict_size=-1, trg_dict_size=-1, lang="en", download=True):
        _need(data_file, "WMT16")
        assert mode in ("train", "test", "val") and lang in ("en", "de")
        self.lang = lang
        with tarfile.open(data_file) as tf:
            def lines(split):
                return [l.split("\t") for l in tf.extractfile(f"wmt16/{split}").read().decode().splitlines() if "\t" in l]

            train = lines("train")
            cur = train if mode == "train" else lines(mode)
        sc, tc = (0, 1) if lang == "en" else (1, 0)
        self.src_dict = self._build([p[sc] for p in train], src_dict_size)
        self.trg_dict = self._build([p[tc] for p in train], trg_dict_size)
        s_unk, t_unk = self.src_dict[self.UNK], self.trg_dict[self.UNK]
        self.data = []
        for p in cur:
            src = [self.src_dict[self.START]] + [self.src_dict.get(w, s_unk) for w in p[sc].split()] + [self.src_dict[self.END]]
            trg = [self.trg_dict.get(w, t_unk) for w in p[tc].split()]
            self.data.append((src, [self.trg_dict[self.START], *trg], [*trg, self.trg_dict[self.END]]))

    def _build(self, sentences, size):
        freq = {}
        for s_ in sentences:
            for w in s_.split():
                freq[w] = freq.get(w, 0) + 1
        words = [self.START, self.END, self.UNK] + [w for w, _ in sorted(freq.items(), key=lambda kv: (-kv[1], kv[0]))]
        if size > 0:
            words = words[:size]
        return {w: i for i, w in enumerate(words)}

    def get_dict(self, lang, reverse=False):
        d = self.src_dict if lang == self.lang else self.trg_dict
        return {v: k for k, v in d.items()} if reverse else d

    def __getitem__(self, idx):
        return tuple(np.array(d) for d in self.data[idx])

    def __len__(self):
        return len(self.data)


__all__ = ["Conll05st", "Imdb", "Imikolov", "Movielens", "UCIHousing", "WMT14", "WMT16", "ViterbiDecoder", "viterbi_decode"]
