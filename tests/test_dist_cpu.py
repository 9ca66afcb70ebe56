"""world_size-2 gloo tests (CPU) of the sharded-search host logic: shard planning at session boundaries,
id offsets, all-gather layout and merge order.  The per-shard search and the merge are played by the CPU
oracle here (the CUDA versions are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import make_iid, make_segments, make_session_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleShard:
    def __init__(self, so, db, seg, id_offset):
        self.so, self.db, self.seg, self.id_offset = so, db, seg, id_offset
        self.ntotal = db.shape[0]
        self.device = None

    def search(self, x, k):
        x = x.numpy() if isinstance(x, torch.Tensor) else x
        return self.so.search_flat(self.db, x, k, seg_off=self.seg, reduce=self.so.REDUCE_MAX, id_offset=self.id_offset)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import search_oracle as so
    from sessionsimilaritysearch_b200.dist import ShardedIndex, shard_bounds
    seg = make_segments(6000, 31)
    db = so.normalize(make_session_rows(seg, 32, 32), so.NORM_UTIL)
    q = so.normalize(make_iid(9, 32, 33), so.NORM_UTIL)
    rows_b, seg_b = shard_bounds(seg, world)
    lo, hi = int(rows_b[rank]), int(rows_b[rank + 1])
    local_seg = seg[seg_b[rank]:seg_b[rank + 1] + 1] - lo
    shard = OracleShard(so, db[lo:hi], local_seg, int(seg_b[rank]))

    def merge(cD, cI, metric):
        D, I = so.topk_merge(cD.numpy(), cI.numpy(), metric)
        return torch.from_numpy(D), torch.from_numpy(I)

    sh = ShardedIndex(shard, merge_fn=merge)
    assert sh.ntotal == db.shape[0]
    D, I = sh.search(q, 20)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), D=np.asarray(D), I=np.asarray(I))
    dist.destroy_process_group()


def test_shard_bounds_cut_at_session_boundaries():
    from sessionsimilaritysearch_b200.dist import shard_bounds
    seg = make_segments(10000, 30)
    for world in (1, 2, 4, 8):
        rows_b, seg_b = shard_bounds(seg, world)
        assert rows_b[0] == 0 and rows_b[-1] == seg[-1] and len(rows_b) == world + 1
        assert np.all(np.diff(rows_b) >= 0) and set(rows_b) <= set(seg)
        assert np.max(np.abs(np.diff(rows_b) - seg[-1] / world)) < 64  # balanced within a few sessions
    # more ranks than sessions: empty shards are legal
    rows_b, seg_b = shard_bounds(np.array([0, 5, 9]), 4)
    assert rows_b[-1] == 9 and len(rows_b) == 5


def test_sharded_search_world2_matches_single(tmp_path, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    seg = make_segments(6000, 31)
    db = oracle.normalize(make_session_rows(seg, 32, 32), oracle.NORM_UTIL)
    q = oracle.normalize(make_iid(9, 32, 33), oracle.NORM_UTIL)
    Do, Io = oracle.search_flat(db, q, 20, seg_off=seg, reduce=oracle.REDUCE_MAX)
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(z["I"], Io) and np.array_equal(z["D"], Do)
