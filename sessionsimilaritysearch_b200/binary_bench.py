"""Benchmark leg of the binary-hash retrieval (what fine_tune_ours.test() executes as committed:
code_len 250 -> 256-bit codes -> faiss.IndexBinaryFlat, fine_tune_ours.py:826,839-843,871-876).  Called by
bench.py --configs ...,binary; returns one dict."""
import numpy as np


def run(env, a, rows=None, nq_list=(1000, 128, 1), k=100):
    import sessionsimilaritysearch_b200 as sss
    torch = env.torch
    rows = int(rows or a.rows_100m)
    nbits, nbytes = 256, 32
    g = torch.Generator(device=env.dev).manual_seed(7)
    ix = sss.IndexBinaryFlat(nbits, device=env.local_rank)
    pool = None
    for lo in range(0, rows, 4_000_000):
        n = min(4_000_000, rows - lo)
        c = torch.randint(0, 256, (n, nbytes), generator=g, device=env.dev, dtype=torch.uint8)
        if pool is None:
            pool = c[:4096].clone()
        ix.add(c)
        del c
    out = {"workload": "%d codes x %d bit (IndexBinaryFlat), top-%d, queries = database codes with 8%% of the bits "
                       "flipped" % (rows, nbits, k), "rows": rows}
    peaks_hbm = None
    for nq in nq_list:
        flip = (torch.rand((nq, nbytes, 8), generator=g, device=env.dev) < 0.08)
        w = (2 ** torch.arange(7, -1, -1, device=env.dev)).to(torch.int32)
        q = pool[:nq] ^ (flip.to(torch.int32) * w).sum(-1).to(torch.uint8)
        for _ in range(3):
            D, I = ix.search(q, k)
        ms = env.timed(lambda i: ix.search(q, k), 5) / 5
        st = ix.stats()
        ix.set_profiling(True)
        ix.search(q, k)
        sp = ix.stats()
        ix.set_profiling(False)
        tensor = st["scan_variant"] in ("ts", "2cta")
        row_bytes = 256 if tensor else nbytes
        blk = {"ms_per_search": ms, "queries_per_s": nq / (ms * 1e-3), "scan_variant": st["scan_variant"],
               "waves": st["waves"], "graph_replay": bool(st["graph"]),
               "scan_kernel_ms": sp["scan_ns"] * 1e-6,
               "db_stream_gbs_as_stored": rows * row_bytes / (ms * 1e-3) / 1e9,
               "db_stream_gbs_packed_equivalent": rows * nbytes / (ms * 1e-3) / 1e9,
               "pairs_per_s": nq * rows / (ms * 1e-3),
               "tensor_tflops_fp8": (2.0 * nq * rows * nbits / (ms * 1e-3) / 1e12) if tensor else None}
        # parity on a sample: the popcount oracle over the first 2M codes against a 2M-code index
        out["nq%d" % nq] = blk
    if env.rank == 0 and not a.no_parity:
        from oracle import search_oracle as so
        n_s = min(rows, 2_000_000)
        gs = torch.Generator(device=env.dev).manual_seed(7)
        c = torch.randint(0, 256, (min(4_000_000, rows), nbytes), generator=gs, device=env.dev, dtype=torch.uint8)[:n_s]
        small = sss.IndexBinaryFlat(nbits, device=env.local_rank)
        small.add(c)
        q = pool[:64] ^ 1
        Ds, Is = small.search(q, k)
        Do, Io = so.search_hamming(c.cpu().numpy(), q.cpu().numpy(), k)
        out["parity"] = {"ok": bool(np.array_equal(Ds.cpu().numpy(), Do) and np.array_equal(Is.cpu().numpy(), Io)),
                         "checked": "64 queries x %d codes against the popcount oracle, distances and ids" % n_s}
    return out
