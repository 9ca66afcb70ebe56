"""Synthetic Amazon-filtered-shaped sessions: the private data and the QAEA text model
are unavailable, so sessions, a stand-in tokenizer and per-string text features are generated from seeds."""
import hashlib

import numpy as np
import torch

ASIN_NUM = 391572          # fine_tune_ours.py:163
ITEM_TYPES = ("c", "ca", "p")


def make_sessions(n, seed, max_len=19):
    """n sessions of action tuples (ts, type, keyword, asin, ptype, brand, title, item_id); length
    2 + Poisson(6) clipped to [2, max_len]; an action is a search with p = 0.3; item ids Zipf(1.1)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        L = int(np.clip(2 + rng.poisson(6), 2, max_len))
        seq = []
        for t in range(L):
            if rng.random() < 0.3:
                kw = "query %d" % int(rng.integers(0, 5000))
                seq.append((t, 's', kw if rng.random() > 0.02 else None, None, None, None, None, 0))
            else:
                item = int(min(rng.zipf(1.1), ASIN_NUM - 1))
                seq.append((t, ITEM_TYPES[int(rng.integers(0, 3))], None, "B%09d" % item, "ptype%d" % (item % 37),
                            "brand%d" % (item % 101), "title of item %d" % item if item % 53 else None, item))
        out.append(seq)
    return out


def split_session(seq, rng):
    """(prefix, suffix) like the us-filtered-split-* files (test_amazon_filterd.py:466,485,546)"""
    cut = int(rng.integers(1, len(seq)))
    return seq[:cut], seq[cut:]


class HashTokenizer:
    """stand-in for AutoTokenizer.from_pretrained('./SavedModel/QAEA') (test_amazon_filterd.py:482): same call
    signature and output keys; token ids are a hash of the words, so equal strings give equal tokens."""

    def __init__(self, vocab=30522):
        self.vocab = vocab

    def __call__(self, texts, padding='max_length', max_length=20, truncation=True, return_tensors="pt"):
        ids = torch.zeros((len(texts), max_length), dtype=torch.long)
        am = torch.zeros((len(texts), max_length), dtype=torch.long)
        for r, t in enumerate(texts):
            words = ["[CLS]"] + (t or "").split()[:max_length - 2] + ["[SEP]"]
            for c, w in enumerate(words):
                ids[r, c] = int.from_bytes(hashlib.blake2s(w.encode(), digest_size=4).digest(), "little") % self.vocab
                am[r, c] = 1
        return {"input_ids": ids, "token_type_ids": torch.zeros_like(ids), "attention_mask": am}


def text_features(token_ids, dim=768):
    """[N, L] token ids -> [N, dim] N(0,1) features, a pure function of the token row (stands in for the frozen
    text encoder + masked mean of model/NodeEmbedding.py:112-125, which is a pure function of the text)"""
    out = torch.empty((token_ids.shape[0], dim), dtype=torch.float32)
    for r in range(token_ids.shape[0]):
        h = hashlib.blake2s(token_ids[r].numpy().tobytes(), digest_size=8).digest()
        g = torch.Generator().manual_seed(int.from_bytes(h, "little") % (2 ** 63))
        out[r] = torch.randn(dim, generator=g)
    return out
