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


class ShardedIndex:
    """index.search() over a row-sharded database.  `inner` is this rank's shard index (built with
    id_offset = the shard's first global row / session id)."""

    def __init__(self, inner, world_size=None, rank=None, group=None, merge_fn=None, metric=_lib.METRIC_IP):
        import torch.distributed as dist
        self.inner = inner
        self.group = group
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.rank = dist.get_rank(group) if rank is None else rank
        self.merge_fn = cuda_merge if merge_fn is None else merge_fn
        self.metric = metric

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

    def search(self, x, k, **kw):
        import torch
        import torch.distributed as dist
        host_in = not type(x).__module__.startswith("torch")
        if host_in and dist.get_backend(self.group) == "nccl":
            xq = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda(self.inner.device, non_blocking=True)
        else:
            xq = x
        D, I = self.inner.search(xq, k, **kw)
        if not type(D).__module__.startswith("torch"):
            D, I = torch.from_numpy(D), torch.from_numpy(I)
        nq, kk = D.shape
        # ONE all-gather: ids and score bits packed side by side as int64 [nq, 2k] (16 B per candidate)
        packed = torch.empty((nq, 2 * kk), dtype=torch.int64, device=D.device)
        packed[:, :kk] = I
        packed[:, kk:] = D.contiguous().view(torch.int32).to(torch.int64)
        gathered = torch.empty((self.world_size * nq, 2 * kk), dtype=torch.int64, device=D.device)
        dist.all_gather_into_tensor(gathered, packed, group=self.group)  # rank-major concatenation
        gathered = gathered.view(self.world_size, nq, 2 * kk)
        cI = gathered[..., :kk].contiguous()
        cD = gathered[..., kk:].to(torch.int32).contiguous().view(torch.float32)
        Dm, Im = self.merge_fn(cD, cI, self.metric)
        if host_in and type(Dm).__module__.startswith("torch"):
            return Dm.cpu().numpy(), Im.cpu().numpy()
        return Dm, Im
