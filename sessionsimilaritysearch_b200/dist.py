"""Row-sharded search across the GPUs of one box (nothing like it exists in the reference).

One process per GPU.  The database rows are partitioned contiguously at session boundaries, every rank
holds one shard in its own index (ids carry the shard's global offset), the query batch is replicated, each
rank produces a local top-k and ONE all-gather of the (score, id) candidates (nq*k*12 bytes per rank) is
followed by the same deterministic (score desc, id asc) k-way merge on every rank.
"""
import ctypes

import numpy as np

from . import _lib


def shard_bounds(seg_off, world_size):
    """Cut [0, n_rows) into world_size contiguous row ranges at session boundaries, balancing rows.
    Returns (row_bounds[world+1], seg_bounds[world+1])."""
    seg_off = np.asarray(seg_off, dtype=np.int64)
    n_seg = len(seg_off) - 1
    n_rows = int(seg_off[-1])
    seg_b = [0]
    for r in range(1, world_size):
        target = n_rows * r // world_size
        s = int(np.searchsorted(seg_off, target, side="left"))
        s = min(max(s, seg_b[-1]), n_seg)
        seg_b.append(s)
    seg_b.append(n_seg)
    seg_b = np.asarray(seg_b, dtype=np.int64)
    return seg_off[seg_b], seg_b


def cuda_merge(cand_D, cand_I, metric=_lib.METRIC_IP):
    """sss_topk_merge on torch CUDA tensors [n_shards, nq, k]"""
    import torch
    lib = _lib.load()
    ns, nq, k = cand_D.shape
    dev = cand_D.device.index
    D = torch.empty((nq, k), dtype=torch.float32, device=cand_D.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=cand_D.device)
    _lib.check(lib.sss_topk_merge(cand_D.contiguous().data_ptr(), cand_I.contiguous().data_ptr(), ns, nq, k, metric,
                                  D.data_ptr(), I.data_ptr(), dev, _lib.current_stream(dev)))
    return D, I


def cuda_merge_packed(gathered, n_shards, nq, k, metric=_lib.METRIC_IP, out=None):
    """sss_topk_merge_packed: `gathered` is the all-gathered uint8 tensor of n_shards packed candidate blocks.
    Returns (D, I, status): status is a device int32 [1], the OR of the shards' status words."""
    import torch
    lib = _lib.load()
    dev = gathered.device.index
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.float32, device=gathered.device),
               torch.empty((nq, k), dtype=torch.int64, device=gathered.device),
               torch.empty(1, dtype=torch.int32, device=gathered.device))
    D, I, status = out
    _lib.check(lib.sss_topk_merge_packed(gathered.data_ptr(), n_shards, nq, k, metric, D.data_ptr(), I.data_ptr(),
                                         status.data_ptr(), dev, _lib.current_stream(dev)))
    return D, I, status


def packed_bytes(nq, k):
    return (nq * k * 12 + 15) // 16 * 16 + 16


def pack_candidates(D, I):
    """host-side restatement of the packed block layout (CPU tests of the sharded logic): uint8 [packed_bytes]"""
    import torch
    nq, k = D.shape
    out = torch.zeros(packed_bytes(nq, k), dtype=torch.uint8)
    out[:nq * k * 8] = I.contiguous().view(-1).view(torch.uint8)
    out[nq * k * 8:nq * k * 12] = D.contiguous().view(-1).view(torch.uint8)
    return out


def unpack_candidates(gathered, n_shards, nq, k):
    """inverse of the packed layout for n_shards blocks laid end to end -> (D [ns, nq, k], I [ns, nq, k])"""
    import torch
    g = gathered.view(n_shards, packed_bytes(nq, k))
    I = g[:, :nq * k * 8].contiguous().view(torch.int64).view(n_shards, nq, k)
    D = g[:, nq * k * 8:nq * k * 12].contiguous().view(torch.float32).view(n_shards, nq, k)
    return D, I


class ShardedIndex:
    """index.search() over a row-sharded database.  `inner` is this rank's shard index (built with
    id_offset = the shard's first global row / session id).

    CUDA path (NCCL): three calls per search and no eager tensor arithmetic — the shard search emits its candidates
    as one packed block, ONE all-gather moves the blocks, the merge kernel reads the gathered blocks in place.
    `merge_fn` (CPU tests over gloo) replaces the CUDA search + merge with host stand-ins on the same layout."""

    def __init__(self, inner, world_size=None, rank=None, group=None, merge_fn=None, metric=_lib.METRIC_IP):
        import torch.distributed as dist
        self.inner = inner
        self.group = group
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.rank = dist.get_rank(group) if rank is None else rank
        self.merge_fn = merge_fn
        self.metric = metric
        self._gathered = None
        self._mine = None
        self._status = None

    @property
    def ntotal(self):
        import torch
        import torch.distributed as dist
        t = torch.tensor([self.inner.ntotal], dtype=torch.int64)
        dev = getattr(self.inner, "device", None)
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda(dev)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def search(self, x, k, out=None, **kw):
        import torch
        import torch.distributed as dist
        host_in = not type(x).__module__.startswith("torch")
        k = int(k)
        if self.merge_fn is not None:  # host stand-ins (gloo): same packed layout, same single all-gather
            D, I = self.inner.search(x, k, **kw)
            if not type(D).__module__.startswith("torch"):
                D, I = torch.from_numpy(np.ascontiguousarray(D)), torch.from_numpy(np.ascontiguousarray(I))
            nq = D.shape[0]
            mine = pack_candidates(D.cpu(), I.cpu())
            gathered = torch.empty(self.world_size * mine.numel(), dtype=torch.uint8)
            dist.all_gather_into_tensor(gathered, mine, group=self.group)
            cD, cI = unpack_candidates(gathered, self.world_size, nq, k)
            Dm, Im = self.merge_fn(cD, cI, self.metric)
            if host_in and type(Dm).__module__.startswith("torch"):
                return Dm.numpy(), Im.numpy()
            return Dm, Im
        # asynchronous pipeline: shard search (captured graph) -> ONE all-gather -> merge, all enqueued back to back;
        # the only host round trip is the status word at the very end
        nq = x.shape[0]
        n = packed_bytes(nq, k)
        if self._mine is None or self._mine.numel() != n:
            import torch as _t
            dev = _t.device("cuda", self.inner.device)
            self._mine = _t.empty(n, dtype=_t.uint8, device=dev)
            self._gathered = _t.empty(self.world_size * n, dtype=_t.uint8, device=dev)
        self.inner.search_packed(x, k, out=self._mine, asynchronous=True, **kw)
        dist.all_gather_into_tensor(self._gathered, self._mine, group=self.group)  # rank-major concatenation
        if self._status is None:
            self._status = torch.empty(1, dtype=torch.int32, device=self._mine.device)
        mo = None if (out is None or host_in) else (out[0], out[1], self._status)
        Dm, Im, status = cuda_merge_packed(self._gathered, self.world_size, nq, k, self.metric, out=mo)
        if host_in:
            Dh, Ih = Dm.cpu(), Im.cpu()   # (synchronises: the status word is final by then)
        if int(status.item()) != 0:
            # some shard overflowed its candidate lists (every rank sees the same OR-ed word): repeat with the
            # synchronous form, whose driver re-runs the shard search with a safer schedule
            self.inner.search_packed(x, k, out=self._mine, asynchronous=False, **kw)
            dist.all_gather_into_tensor(self._gathered, self._mine, group=self.group)
            Dm, Im, status = cuda_merge_packed(self._gathered, self.world_size, nq, k, self.metric, out=mo)
            if host_in:
                Dh, Ih = Dm.cpu(), Im.cpu()
        if host_in:
            return Dh.numpy(), Ih.numpy()
        return Dm, Im
