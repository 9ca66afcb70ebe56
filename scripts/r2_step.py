"""One search step of the headline workload for ncu launch lists: python scripts/r2_step.py [nq] [n_steps] [rows]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sessionsimilaritysearch_b200 as sss  # noqa: E402


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
    sys.argv = sys.argv[:1]
    a = bench.parse()
    a.nq = nq
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
    for _ in range(n_steps):
        ix.search(q, a.k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ix.search(q, a.k)
    e1.record()
    torch.cuda.synchronize()
    print("nq", nq, "ms/step", e0.elapsed_time(e1) / 10, ix.stats())


if __name__ == "__main__":
    main()
