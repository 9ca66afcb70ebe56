"""How much of the headline step is host time?  Synchronous searches (one stream sync + status read per call) against
the same captured graph replayed back to back through the asynchronous packed form (one sync at the end):
python scripts/r2_host_overhead.py [nq] [rows]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sessionsimilaritysearch_b200 as sss  # noqa: E402


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    sys.argv = sys.argv[:1]
    a = bench.parse()
    a.nq, a.rows = nq, rows
    env = bench.Env()
    lens = bench.session_lengths(rows, 1234)
    ix = sss.IndexFlatIP(a.d, device=0, mode="exact")
    pool = []
    for rows_c, _, bases in bench.shard_chunks(env, a, lens, 4321):
        ix.add(rows_c, norm=sss.NORM_UTIL)
        pool.append(bases.clone())
    ix.set_segments(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), "max")
    pool = torch.cat(pool)
    g = torch.Generator(device="cuda").manual_seed(99)
    pick = torch.randint(0, pool.shape[0], (nq,), generator=g, device="cuda")
    q = sss.normalize(pool[pick] + 0.3 * torch.randn((nq, a.d), generator=g, device="cuda"))
    n = 30
    for _ in range(5):
        ix.search(q, a.k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        ix.search(q, a.k)
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    sync_ms = e0.elapsed_time(e1) / n
    out = None
    for _ in range(3):
        out = ix.search_packed(q, a.k, out=out, asynchronous=True)
    torch.cuda.synchronize()
    e0.record()
    t2 = time.perf_counter()
    for _ in range(n):
        out = ix.search_packed(q, a.k, out=out, asynchronous=True)
    t3 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    async_ms = e0.elapsed_time(e1) / n
    print("nq %d rows %d: synchronous %.4f ms/step (host wall %.4f), graph replays back to back %.4f ms/step "
          "(host enqueue %.4f ms/call) -> host-exposed %.1f us per synchronous step"
          % (nq, rows, sync_ms, (t1 - t0) * 1e3 / n, async_ms, (t3 - t2) * 1e3 / n, (sync_ms - async_ms) * 1e3))


if __name__ == "__main__":
    main()
