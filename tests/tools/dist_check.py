"""torchrun worker: sharded search == single-index search == oracle (run by tests/test_dist_gpu.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sessionsimilaritysearch_b200 as sss  # noqa: E402
from conftest import make_iid, make_segments, make_session_rows  # noqa: E402
from oracle import search_oracle as so  # noqa: E402
from sessionsimilaritysearch_b200.dist import ShardedIndex, shard_bounds  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seg = make_segments(300000, 41)
    db = make_session_rows(seg, 128, 42)
    q = make_iid(300, 128, 43)
    for reduce in ("max", None):
        rows_b, seg_b = shard_bounds(seg, world)
        lo, hi = int(rows_b[rank]), int(rows_b[rank + 1])
        ix = sss.IndexFlatIP(128, device=local, id_offset=int(seg_b[rank]) if reduce else lo)
        ix.add(db[lo:hi], norm=sss.NORM_UTIL)
        if reduce:
            ix.set_segments(seg[seg_b[rank]:seg_b[rank + 1] + 1] - lo, reduce)
        sh = ShardedIndex(ix)
        assert sh.ntotal == db.shape[0]
        qn = sss.normalize(q, device=local)
        D, I = sh.search(qn, 100)                                   # host queries in, host results out
        Dd, Id = sh.search(torch.from_numpy(qn).cuda(local), 100)   # device in, device out
        assert np.array_equal(Id.cpu().numpy(), I) and np.array_equal(Dd.cpu().numpy(), D)
        if rank == 0:
            Do, Io = so.search_flat(so.normalize(db, 1), so.normalize(q, 1), 100, seg_off=seg if reduce else None,
                                    reduce=so.REDUCE_MAX if reduce else so.REDUCE_NONE)
            assert np.array_equal(I, Io), "sharded ids differ from the oracle (reduce=%s)" % reduce
            assert np.array_equal(D.view(np.uint32), Do.view(np.uint32)), "sharded scores differ"
        dist.barrier()
    if rank == 0:
        print("DIST_CHECK_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
