"""Flat and binary indexes with the call surface the reference uses from faiss.

    build_index(emb, metric)                 test_amazon_filterd.py:207-223
    index.add(x); index.search(x, K)         test_amazon_filterd.py:214,578; fine_tune_ours.py:849,882
    IndexBinaryFlat(nbits).add/search        fine_tune_ours.py:839-843,871-876

numpy in -> numpy out (host buffers, H2D/D2H inside the call, like faiss' CPU API);
torch CUDA tensors in -> torch CUDA tensors out (nothing leaves HBM).
All arithmetic runs in libsss_b200.so (CUDA, sm_100a); there is no CPU path.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (METRIC_IP, METRIC_L2, MODES, NORM_FT, NORM_NONE, NORM_TORCH, NORM_UTIL, REDUCES, check)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _f32_host(x, d=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim != 2 or (d is not None and x.shape[1] != d):
        raise ValueError("expected a float32 array of shape [n, %s], got %s" % (d, x.shape))
    return x


def _f32_dev(x, d, device):
    import torch
    if x.dim() != 2 or (d is not None and x.shape[1] != d):
        raise ValueError("expected a tensor of shape [n, %s], got %s" % (d, tuple(x.shape)))
    if x.device.type != "cuda" or x.device.index != device:
        raise ValueError("tensor must live on cuda:%d, got %s" % (device, x.device))
    return x.detach().to(torch.float32).contiguous()


class _FlatIndex:
    """Common body of IndexFlatIP / IndexFlatL2."""
    _metric = METRIC_IP

    def __init__(self, d, device=None, id_offset=0, mode="exact"):
        self._lib = _lib.load()
        self.d = int(d)
        self.device = _lib.current_device() if device is None else int(device)
        self.mode = mode
        self.id_offset = int(id_offset)
        h = ctypes.c_void_p()
        check(self._lib.sss_index_create(ctypes.byref(h), self.device, self.d, self._metric, self.id_offset))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.sss_index_destroy(h)

    @property
    def ntotal(self):
        return int(self._lib.sss_index_ntotal(self._h))

    def add(self, x, norm=NORM_NONE):
        """index.add(x).  `norm` fuses the reference's normalize() into the GPU add
        (NORM_UTIL: util_amazon_filtered.py:28-31, NORM_FT: fine_tune_ours.py:38-40)."""
        if _is_torch(x):
            x = _f32_dev(x, self.d, self.device)
            check(self._lib.sss_index_add(self._h, x.data_ptr(), x.shape[0], 1, norm,
                                          _lib.current_stream(self.device)))
            import torch
            torch.cuda.current_stream(self.device).synchronize()  # x may be freed by the caller
        else:
            x = _f32_host(x, self.d)
            check(self._lib.sss_index_add(self._h, x.ctypes.data, x.shape[0], 0, norm,
                                          _lib.current_stream(self.device)))

    def set_segments(self, seg_off, reduce="max"):
        """Declare contiguous sessions: rows [seg_off[s], seg_off[s+1]) are the subsessions of session s.
        search() then returns session ids scored by the max (or sum) over their rows."""
        if reduce not in REDUCES:
            raise ValueError("reduce must be one of None, 'max', 'sum'")
        st = _lib.current_stream(self.device)
        if REDUCES[reduce] == 0:
            check(self._lib.sss_index_set_segments(self._h, None, 0, 0, st))
            return
        so = np.ascontiguousarray(np.asarray(seg_off.cpu() if _is_torch(seg_off) else seg_off), dtype=np.int64)
        check(self._lib.sss_index_set_segments(self._h, so.ctypes.data, so.shape[0] - 1, REDUCES[reduce], st))

    def search(self, x, k, mode=None, out=None):
        """D, I = index.search(x, K): D float32 [nq, K] best first, I int64 [nq, K]; ties -> smaller id.
        out=(D, I): contiguous CUDA tensors of those shapes to write into (device queries only) — a loop over query
        batches then allocates nothing (a fresh cudaMalloc by the caching allocator costs up to 100 ms on some hosts)."""
        m = MODES[self.mode if mode is None else mode]
        k = int(k)
        if _is_torch(x):
            import torch
            x = _f32_dev(x, self.d, self.device)
            nq = x.shape[0]
            if out is not None:
                D, I = out
                if not (D.is_cuda and I.is_cuda and D.dtype == torch.float32 and I.dtype == torch.int64 and
                        tuple(D.shape) == (nq, k) and tuple(I.shape) == (nq, k) and D.is_contiguous() and
                        I.is_contiguous() and D.device == x.device and I.device == x.device):
                    raise ValueError("out must be contiguous CUDA tensors (float32 [nq, k], int64 [nq, k]) on the index's device")
            else:
                D = torch.empty((nq, k), dtype=torch.float32, device=x.device)
                I = torch.empty((nq, k), dtype=torch.int64, device=x.device)
            check(self._lib.sss_index_search(self._h, x.data_ptr(), nq, k, m, 1, D.data_ptr(), I.data_ptr(), 1,
                                             _lib.current_stream(self.device)))
            return D, I
        x = _f32_host(x, self.d)
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        check(self._lib.sss_index_search(self._h, x.ctypes.data, nq, k, m, 0, D.ctypes.data, I.ctypes.data, 0,
                                         _lib.current_stream(self.device)))
        return D, I

    def search_packed(self, x, k, mode=None, out=None, asynchronous=False):
        """The sharded path's form of search(): one uint8 CUDA tensor holding this shard's candidates as
        [ids int64 nq*k | scores fp32 nq*k | pad | 16-byte status trailer] (sss_packed_bytes), ready to be all-gathered
        as it is.  asynchronous=True returns as soon as the work is enqueued: the status word (overflow -> the search
        must be repeated synchronously) travels in the trailer and is checked by the caller after the merge."""
        import torch
        m = MODES[self.mode if mode is None else mode]
        k = int(k)
        dev = torch.device("cuda", self.device)
        if _is_torch(x):
            x = _f32_dev(x, self.d, self.device)
            ptr, on_dev, nq = x.data_ptr(), 1, x.shape[0]
        else:
            x = _f32_host(x, self.d)
            ptr, on_dev, nq = x.ctypes.data, 0, x.shape[0]
        n = int(self._lib.sss_packed_bytes(nq, k))
        if out is None or out.numel() != n:
            out = torch.empty(n, dtype=torch.uint8, device=dev)
        check(self._lib.sss_index_search_packed(self._h, ptr, nq, k, m, on_dev, out.data_ptr(), int(bool(asynchronous)),
                                                _lib.current_stream(self.device)))
        return out

    def set_profiling(self, on=True):
        """time every scan-kernel launch with CUDA events on the launching stream (bench.py roofline); a profiled
        search runs as plain launches instead of replaying its captured CUDA graph"""
        check(self._lib.sss_index_set_profiling(self._h, int(bool(on))))

    def stats(self):
        st = self._lib.sss_index_stat
        return {"kernels": int(st(self._h, 0)), "waves": int(st(self._h, 1)), "reruns": int(st(self._h, 2)),
                "scan_ns": int(st(self._h, 3)), "scan_launches": int(st(self._h, 4)),
                "refine_candidates": int(st(self._h, 5)), "refine_rescored": int(st(self._h, 6)),
                "refine_sessions": int(st(self._h, 7)), "refine_calls": int(st(self._h, 8)),
                "overflow_reason": int(st(self._h, 24)), "graph": int(st(self._h, 26)),
                "scan_variant": ("fp32", "ss", "ts", "2cta", "kloop")[int(st(self._h, 25))]}


class IndexFlatIP(_FlatIndex):
    """faiss.IndexFlatIP(d): exact inner-product top-k, descending."""
    _metric = METRIC_IP


class IndexFlatL2(_FlatIndex):
    """faiss.IndexFlatL2(d): exact squared-L2 top-k, ascending."""
    _metric = METRIC_L2


class IndexBinaryFlat:
    """faiss.IndexBinaryFlat(nbits): Hamming top-k over np.packbits codes (fine_tune_ours.py:842-843,876)."""

    def __init__(self, nbits, device=None, id_offset=0):
        self._lib = _lib.load()
        self.d = int(nbits)
        self.code_size = self.d // 8
        self.device = _lib.current_device() if device is None else int(device)
        h = ctypes.c_void_p()
        check(self._lib.sss_binary_create(ctypes.byref(h), self.device, self.d, int(id_offset)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.sss_binary_destroy(h)

    @property
    def ntotal(self):
        return int(self._lib.sss_binary_ntotal(self._h))

    def set_profiling(self, on=True):
        check(self._lib.sss_binary_set_profiling(self._h, int(bool(on))))

    def stats(self):
        st = self._lib.sss_binary_stat
        return {"kernels": int(st(self._h, 0)), "waves": int(st(self._h, 1)), "reruns": int(st(self._h, 2)),
                "scan_ns": int(st(self._h, 3)), "scan_launches": int(st(self._h, 4)),
                "overflow_reason": int(st(self._h, 24)), "graph": int(st(self._h, 26)),
                "scan_variant": ("popcount", "ss", "ts", "2cta", "kloop")[int(st(self._h, 25))]}

    def _codes(self, x):
        if _is_torch(x):
            import torch
            if x.dtype != torch.uint8 or x.dim() != 2 or x.shape[1] != self.code_size:
                raise ValueError("expected uint8 codes of shape [n, %d]" % self.code_size)
            return x.contiguous(), True
        x = np.ascontiguousarray(x)
        if x.dtype != np.uint8 or x.ndim != 2 or x.shape[1] != self.code_size:
            raise ValueError("expected uint8 codes of shape [n, %d]" % self.code_size)
        return x, False

    def add(self, x):
        x, dev = self._codes(x)
        ptr = x.data_ptr() if dev else x.ctypes.data
        check(self._lib.sss_binary_add(self._h, ptr, x.shape[0], int(dev), _lib.current_stream(self.device)))
        if dev:
            import torch
            torch.cuda.current_stream(self.device).synchronize()

    def search(self, x, k):
        x, dev = self._codes(x)
        nq, k = x.shape[0], int(k)
        if dev:
            import torch
            D = torch.empty((nq, k), dtype=torch.int32, device=x.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=x.device)
            check(self._lib.sss_binary_search(self._h, x.data_ptr(), nq, k, 1, D.data_ptr(), I.data_ptr(), 1,
                                              _lib.current_stream(self.device)))
            return D, I
        D = np.empty((nq, k), dtype=np.int32)
        I = np.empty((nq, k), dtype=np.int64)
        check(self._lib.sss_binary_search(self._h, x.ctypes.data, nq, k, 0, D.ctypes.data, I.ctypes.data, 0,
                                          _lib.current_stream(self.device)))
        return D, I


def normalize(vec, mode=NORM_UTIL, device=None):
    """normalize(vec) of util_amazon_filtered.py:28-31 (mode=NORM_UTIL, 1-D or 2-D) and of
    fine_tune_ours.py:38-40 (mode=NORM_FT, 2-D), computed on the GPU in a fixed summation order."""
    lib = _lib.load()
    if _is_torch(vec):
        import torch
        dev = vec.device.index
        x = vec.detach().to(torch.float32).contiguous()
        x2 = x.view(1, -1) if x.dim() == 1 else x
        out = torch.empty_like(x2)
        check(lib.sss_normalize(x2.data_ptr(), out.data_ptr(), x2.shape[0], x2.shape[1], mode, 1, dev,
                                _lib.current_stream(dev)))
        return out.view(x.shape)
    dev = _lib.current_device() if device is None else device
    x = np.ascontiguousarray(vec, dtype=np.float32)
    x2 = x.reshape(1, -1) if x.ndim == 1 else x
    out = np.empty_like(x2)
    check(lib.sss_normalize(x2.ctypes.data, out.ctypes.data, x2.shape[0], x2.shape[1], mode, 0, dev,
                            _lib.current_stream(dev)))
    return out.reshape(x.shape)


def pack_sign_bits(x, device=None):
    """np.packbits(((x + 1) / 2).astype(int), axis=1) for BinarizeHead outputs in {-1, 0, +1}
    (fine_tune_ours.py:839-840 on model/model.py:137)."""
    lib = _lib.load()
    if _is_torch(x):
        import torch
        dev = x.device.index
        x = x.detach().to(torch.float32).contiguous()
        out = torch.empty((x.shape[0], (x.shape[1] + 7) // 8), dtype=torch.uint8, device=x.device)
        check(lib.sss_pack_sign_bits(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], 1, dev,
                                     _lib.current_stream(dev)))
        return out
    dev = _lib.current_device() if device is None else device
    x = _f32_host(x)
    out = np.empty((x.shape[0], (x.shape[1] + 7) // 8), dtype=np.uint8)
    check(lib.sss_pack_sign_bits(x.ctypes.data, out.ctypes.data, x.shape[0], x.shape[1], 0, dev,
                                 _lib.current_stream(dev)))
    return out


def build_index(emb, metric, device=None, mode="exact"):
    """build_index(emb, metric) of test_amazon_filterd.py:207-223: 'cos' -> inner product over
    normalize(emb) (util_amazon_filtered.py:28-31, fused into the GPU add), 'l2', 'ip'."""
    print(emb.shape)
    if metric == 'cos':
        index = IndexFlatIP(emb.shape[1], device=device, mode=mode)
        index.add(emb, norm=NORM_UTIL)
    elif metric == 'l2':
        index = IndexFlatL2(emb.shape[1], device=device, mode=mode)
        index.add(emb)
    elif metric == 'ip':
        index = IndexFlatIP(emb.shape[1], device=device, mode=mode)
        index.add(emb)
    else:
        raise RuntimeError("Unregnozed metric", metric)
    return index
