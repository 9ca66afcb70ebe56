"""Evaluation metrics on the retrieved ids: get_score / get_ave_score of the reference (fine_tune_ours.py:42-97,
883-897; test_amazon_filterd.py:669-673), batched over I[nq, K] instead of nq x K Python calls.

    all_jaccard / cur_jaccard / all_product_type_score   -> sss_pair_scores   (CUDA, one thread per pair)
    all_query_score / all_product_title_score            -> sss_seqratio_pairs (native host threads)

A session is a list of action tuples (ts, type, keyword, asin, ptype, brand, title, item_id); `test_data` is a list
of (prefix, suffix) pairs, `train_data` a list of sessions — the shapes the reference passes.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check

SIM_TYPES = ("all_jaccard", "cur_jaccard", "all_query_score", "all_product_title_score", "all_product_type_score")


def _items(session):        # get_item, util_amazon_filtered.py:33-34 (a set: order is irrelevant to Jaccard)
    return sorted(set(a[-1] for a in session if a[1] != 's'))


def _types(session):        # get_item_type, util_amazon_filtered.py:59-60
    return [a[4] for a in session if a[1] != 's' if a[4] is not None]


def _queries(session):      # get_query(pad=False), util_amazon_filtered.py:234-236
    return [a[2] for a in session if a[1] == 's' and a[2] is not None]


def _titles(session):       # get_session_item_title, util_amazon_filtered.py:36-37
    return [a[-2] if a[-2] is not None else '' for a in session if a[1] != 's']


def _csr(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum([len(l) for l in lists], out=off[1:])
    vals = np.fromiter((v for l in lists for v in l), dtype=np.int64, count=int(off[-1]))
    return off, vals


class _StringSeqs:
    """lists of strings -> the flat UTF-32 arrays of sss_string_seqs_t (host)"""

    def __init__(self, lists):
        self.seq_off = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum([len(l) for l in lists], out=self.seq_off[1:])
        strings = [s for l in lists for s in l]
        self.str_off = np.zeros(len(strings) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in strings], out=self.str_off[1:])
        joined = "".join(strings)
        self.chars = np.frombuffer(joined.encode("utf-32-le"), dtype=np.uint32).copy() if joined else np.zeros(1, np.uint32)
        self.c = _lib.StringSeqs(len(lists), self.seq_off.ctypes.data, self.str_off.ctypes.data, self.chars.ctypes.data)


def score_matrix(I, test_data, train_data, sim_type, device=None, n_threads=0):
    """The reference's `gt` (fine_tune_ours.py:883-888): float32 [nq, K], gt[i, j] = get_score(test_data[i],
    (train_data[I[i, j]], []), sim_type)."""
    if sim_type not in SIM_TYPES:
        raise RuntimeError("unrecognized sim type: %s" % sim_type)
    lib = _lib.load()
    I_np = np.ascontiguousarray(I.cpu().numpy() if torch.is_tensor(I) else I, dtype=np.int64)
    nq, k = I_np.shape
    if len(test_data) != nq:
        raise ValueError("I has %d rows for %d test sessions" % (nq, len(test_data)))
    whole = [t[0] + t[1] for t in test_data]
    if sim_type in ("all_query_score", "all_product_title_score"):
        pick = _queries if sim_type == "all_query_score" else _titles
        a, b = _StringSeqs([pick(s) for s in whole]), _StringSeqs([pick(s) for s in train_data])
        out = np.empty((nq, k), dtype=np.float32)
        check(lib.sss_seqratio_pairs(ctypes.byref(a.c), ctypes.byref(b.c), I_np.ctypes.data, nq, k,
                                     1 if sim_type == "all_query_score" else 0, out.ctypes.data, int(n_threads)))
        return out
    dev_i = _lib.current_device() if device is None else int(device)
    dev = torch.device("cuda", dev_i)
    if sim_type == "all_product_type_score":
        vocab = {}
        enc = lambda s: [vocab.setdefault(t, len(vocab)) for t in _types(s)]
        a_lists, b_lists, kind = [enc(s) for s in whole], [enc(s) for s in train_data], 1
    else:
        src = whole if sim_type == "all_jaccard" else [t[0] for t in test_data]
        a_lists, b_lists, kind = [_items(s) for s in src], [_items(s) for s in train_data], 0
    (a_off, a_vals), (b_off, b_vals) = _csr(a_lists), _csr(b_lists)
    t = [torch.from_numpy(x).to(dev) for x in (a_off, a_vals, b_off, b_vals, I_np)]
    out = torch.empty((nq, k), dtype=torch.float32, device=dev)
    check(lib.sss_pair_scores(kind, t[0].data_ptr(), t[1].data_ptr(), nq, t[2].data_ptr(), t[3].data_ptr(),
                              len(train_data), t[4].data_ptr(), k, out.data_ptr(), dev_i, _lib.current_stream(dev_i)))
    return out.cpu().numpy()


def get_ave_score(I, test_data, train_data, sim_type, device=None):
    """get_ave_score(I, test_data, train_data, sim_type) of fine_tune_ours.py:89-96: np.mean of the score matrix"""
    return np.mean(score_matrix(I, test_data, train_data, sim_type, device=device))


def get_score(data_a, data_b, sim_type, device=None):
    """get_score(data_a, data_b, sim_type) of fine_tune_ours.py:42-87 for ONE pair (a batch of one through the same
    native path; use score_matrix / get_ave_score for the nq x K evaluation loop)."""
    gt = score_matrix(np.zeros((1, 1), np.int64), [data_a], [data_b[0] + data_b[1]], sim_type, device=device)
    return float(gt[0, 0])
