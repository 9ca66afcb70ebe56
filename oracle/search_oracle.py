"""numpy/ctypes front end of the CPU oracle for the retrieval path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.

Two oracles:
  O1  the reference's arithmetic as written: numpy `normalize` (util_amazon_filtered.py:28-31,
      fine_tune_ours.py:38-40) -> `Q @ D.T` -> top-k.  Summation order is whatever BLAS does, so O1 is
      compared tie-tolerantly.
  O2  search_oracle.c: fixed-order fp32 FMA restatement, (score desc, id asc).  The CUDA `fp32` and
      `exact` modes must match O2 bit for bit.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsss_oracle.so")
_SRC = os.path.join(_HERE, "search_oracle.c")

NORM_NONE, NORM_UTIL, NORM_FT, NORM_TORCH = 0, 1, 2, 3
METRIC_IP, METRIC_L2 = 0, 1
REDUCE_NONE, REDUCE_MAX, REDUCE_SUM = 0, 1, 2


def build(force=False):
    """gcc the C restatement next to its source (x86-64-v3 for hardware FMA; -ffp-contract=off so that
    only the explicit fmaf calls fuse)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC,
           "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def normalize(x, mode):
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    out = np.empty_like(x)
    lib().o_normalize(_p(x, ctypes.c_float), _p(out, ctypes.c_float), ctypes.c_int64(x.shape[0]),
                      ctypes.c_int(x.shape[1]), ctypes.c_int(mode))
    return out


def search_flat(db, q, k, metric=METRIC_IP, seg_off=None, reduce=REDUCE_NONE, id_offset=0):
    db = np.ascontiguousarray(db, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, d = db.shape
    nq = q.shape[0]
    assert q.shape[1] == d
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    if seg_off is not None:
        seg_off = np.ascontiguousarray(seg_off, dtype=np.int64)
        so, ns = _p(seg_off, ctypes.c_int64), seg_off.shape[0] - 1
    else:
        so, ns = None, 0
    rc = lib().o_search_flat(_p(db, ctypes.c_float), ctypes.c_int64(n), ctypes.c_int(d), _p(q, ctypes.c_float),
                             ctypes.c_int64(nq), ctypes.c_int(k), ctypes.c_int(metric), so, ctypes.c_int64(ns),
                             ctypes.c_int(reduce), ctypes.c_int64(id_offset), _p(D, ctypes.c_float),
                             _p(I, ctypes.c_int64))
    if rc != 0:
        raise RuntimeError("oracle o_search_flat failed: %d" % rc)
    return D, I


def search_hamming(db, q, k, id_offset=0):
    db = np.ascontiguousarray(db, dtype=np.uint8)
    q = np.ascontiguousarray(q, dtype=np.uint8)
    n, nb = db.shape
    nq = q.shape[0]
    D = np.empty((nq, k), dtype=np.int32)
    I = np.empty((nq, k), dtype=np.int64)
    lib().o_search_hamming(_p(db, ctypes.c_uint8), ctypes.c_int64(n), ctypes.c_int(nb), _p(q, ctypes.c_uint8),
                           ctypes.c_int64(nq), ctypes.c_int(k), ctypes.c_int64(id_offset), _p(D, ctypes.c_int32),
                           _p(I, ctypes.c_int64))
    return D, I


def pack_sign_bits(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, nbits = x.shape
    out = np.empty((n, (nbits + 7) // 8), dtype=np.uint8)
    lib().o_pack_sign_bits(_p(x, ctypes.c_float), _p(out, ctypes.c_uint8), ctypes.c_int64(n), ctypes.c_int(nbits))
    return out


def item_vote(D, I, item_off, items, K):
    D = np.ascontiguousarray(D, dtype=np.float32)
    I = np.ascontiguousarray(I, dtype=np.int64)
    item_off = np.ascontiguousarray(item_off, dtype=np.int64)
    items = np.ascontiguousarray(items, dtype=np.int64)
    nq, s = D.shape
    oi = np.empty((nq, K), dtype=np.int64)
    ow = np.empty((nq, K), dtype=np.float32)
    lib().o_item_vote(_p(D, ctypes.c_float), _p(I, ctypes.c_int64), ctypes.c_int64(nq), ctypes.c_int(s),
                      _p(item_off, ctypes.c_int64), _p(items, ctypes.c_int64), ctypes.c_int(K),
                      _p(oi, ctypes.c_int64), _p(ow, ctypes.c_float))
    return oi, ow


def topk_merge(cD, cI, metric=METRIC_IP):
    cD = np.ascontiguousarray(cD, dtype=np.float32)
    cI = np.ascontiguousarray(cI, dtype=np.int64)
    ns, nq, k = cD.shape
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    lib().o_topk_merge(_p(cD, ctypes.c_float), _p(cI, ctypes.c_int64), ctypes.c_int(ns), ctypes.c_int64(nq),
                       ctypes.c_int(k), ctypes.c_int(metric), _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
    return D, I


# ---------------------------------------------------------------------------------------------------
# O1: the reference's arithmetic as written (numpy), used tie-tolerantly and as the CPU baseline.
# ---------------------------------------------------------------------------------------------------

def normalize_util_numpy(vec):
    """util_amazon_filtered.py:28-31, restated."""
    if len(vec.shape) == 1:
        return vec / np.sqrt(np.clip(np.sum(vec ** 2), 1e-6, None))
    return vec / np.sqrt(np.clip(np.sum(vec ** 2, axis=1), 1e-6, None)).reshape(-1, 1)


def normalize_ft_numpy(v):
    """fine_tune_ours.py:38-40, restated."""
    norm = np.linalg.norm(v, axis=1) + 1e-4
    return v / np.expand_dims(norm, -1)


def search_blas(db, q, k, metric=METRIC_IP, seg_off=None, reduce=REDUCE_NONE, chunk=131072, threads=None):
    """faiss-flat-style CPU search: chunked sgemm + top-k + merge.  Uses torch's
    MKL sgemm with all host threads.  Ties -> smaller id.  This is the timed CPU baseline."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    dbt = torch.from_numpy(np.ascontiguousarray(db, dtype=np.float32))
    qt = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
    n = dbt.shape[0]
    nq = qt.shape[0]
    best_s = torch.full((nq, 0), 0.0)
    best_i = torch.zeros((nq, 0), dtype=torch.int64)
    if seg_off is not None and reduce != REDUCE_NONE:
        seg_off_t = torch.from_numpy(np.ascontiguousarray(seg_off, dtype=np.int64))
        n_seg = seg_off_t.numel() - 1
        row_seg = torch.repeat_interleave(torch.arange(n_seg), seg_off_t[1:] - seg_off_t[:-1])
        if reduce == REDUCE_SUM:
            acc = torch.zeros((n_seg, dbt.shape[1]))
            acc.index_add_(0, row_seg, dbt)
            dbt, n, seg_off = acc, n_seg, None
            reduce = REDUCE_NONE
    qn = (qt * qt).sum(1, keepdim=True) if metric == METRIC_L2 else None
    s0 = 0
    while s0 < n:
        e0 = min(n, s0 + chunk)
        if reduce == REDUCE_MAX:
            # cut chunks at segment boundaries
            sg0 = int(row_seg[s0])
            sg1 = int(row_seg[e0 - 1]) + 1 if e0 < n else n_seg
            if e0 < n:
                e0 = int(seg_off_t[sg1 - 1]) if int(seg_off_t[sg1 - 1]) > s0 else int(seg_off_t[sg1])
                sg1 = int(row_seg[e0 - 1]) + 1
        blk = dbt[s0:e0]
        sc = qt @ blk.T
        if metric == METRIC_L2:
            sc = -(qn + (blk * blk).sum(1)[None, :] - 2 * sc)
        if reduce == REDUCE_MAX:
            local = row_seg[s0:e0] - sg0
            red = torch.full((nq, sg1 - sg0), float("-inf"))
            red.scatter_reduce_(1, local[None, :].expand(nq, -1), sc, reduce="amax", include_self=True)
            sc = red
            ids = torch.arange(sg0, sg1)
        else:
            ids = torch.arange(s0, e0)
        kk = min(k, sc.shape[1])
        ts, ti = torch.topk(sc, kk, dim=1)
        best_s = torch.cat([best_s, ts], 1)
        best_i = torch.cat([best_i, ids[ti]], 1)
        if best_s.shape[1] > k:
            # (score desc, id asc): sort by id first, then stable by score
            o = torch.argsort(best_i, dim=1, stable=True)
            best_s, best_i = torch.gather(best_s, 1, o), torch.gather(best_i, 1, o)
            o = torch.argsort(best_s, dim=1, descending=True, stable=True)[:, :k]
            best_s, best_i = torch.gather(best_s, 1, o), torch.gather(best_i, 1, o)
        s0 = e0
    o = torch.argsort(best_i, dim=1, stable=True)
    best_s, best_i = torch.gather(best_s, 1, o), torch.gather(best_i, 1, o)
    o = torch.argsort(best_s, dim=1, descending=True, stable=True)[:, :k]
    best_s, best_i = torch.gather(best_s, 1, o), torch.gather(best_i, 1, o)
    D = best_s.numpy()
    if metric == METRIC_L2:
        D = -D
    return D.astype(np.float32), best_i.numpy()


def search_float64(db, q, k, metric=METRIC_IP):
    """float64 brute force used to sanity check O2 itself (independent of fmaf ordering)."""
    db64 = np.asarray(db, dtype=np.float64)
    q64 = np.asarray(q, dtype=np.float64)
    if metric == METRIC_IP:
        sc = q64 @ db64.T
    else:
        sc = -((q64 ** 2).sum(1)[:, None] + (db64 ** 2).sum(1)[None, :] - 2 * q64 @ db64.T)
    idx = np.argsort(-sc, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(sc, idx, 1), idx
