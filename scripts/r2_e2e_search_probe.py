"""Per-batch timings and driver statistics of the search half of the configs[4] chain (why does it vary run to run?):
python scripts/r2_e2e_search_probe.py [n_sessions]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import encoder_common as ec  # noqa: E402
import sessionsimilaritysearch_b200 as sss  # noqa: E402
from sessionsimilaritysearch_b200 import featurize, pipeline, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    in_dim, hidden, n_layers, out_dim, msl = 768, 800, 3, 1600, 20
    enc = sss.SessionEncoder(ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 11), in_dim=in_dim, hidden=hidden,
                             n_layers=n_layers, out_dim=out_dim, max_seq_len=msl, device=0)
    sessions_all = synth.make_sessions(n, 17)
    vocab = featurize.QueryVocab()
    items = set([0])
    for s in sessions_all:
        for act in s:
            if act[1] == 's':
                vocab(act[2])
            else:
                items.add(act[-1])
    item_ids = np.asarray(sorted(items), dtype=np.int64)
    g = torch.Generator().manual_seed(3)
    cache = featurize.FeatureCache(torch.randn((len(vocab), in_dim), generator=g), item_ids,
                                   torch.randn((len(item_ids), in_dim), generator=g), 0)
    pipe = pipeline.SessionSearchPipeline(enc, cache, vocab)
    pipe.build(sessions_all)
    queries = [s[:max(2, (2 * len(s)) // 3)] for s in sessions_all]
    emb = sss.normalize(pipe.encode_queries(queries))
    torch.cuda.synchronize()
    for rep in range(2):
        rows = []
        t_all = time.perf_counter()
        for lo in range(0, emb.shape[0], 2048):
            t0 = time.perf_counter()
            pipe.index.search(emb[lo:lo + 2048].contiguous(), 100)
            torch.cuda.synchronize()
            st = pipe.index.stats()
            rows.append(((time.perf_counter() - t0) * 1e3, st["reruns"], st["waves"], st["overflow_reason"], st["graph"],
                         st["scan_variant"]))
        print("rep %d: %.3f s total" % (rep, time.perf_counter() - t_all))
        print(" ".join("%.1f/%d/%d/%d/%d" % r[:5] for r in rows), rows[0][5])


    # where a batch's time goes: CUDA events around the scan launches (plain launches, no graph)
    pipe.index.set_profiling(True)
    for lo in (0, 2048 * 10):
        t0 = time.perf_counter()
        pipe.index.search(emb[lo:lo + 2048].contiguous(), 100)
        torch.cuda.synchronize()
        st = pipe.index.stats()
        print("profiled batch at %d: %.2f ms wall, scan kernels %.2f ms in %d launches, %s" % (
            lo, (time.perf_counter() - t0) * 1e3, st["scan_ns"] / 1e6, st["scan_launches"],
            {k: st[k] for k in ("waves", "kernels", "refine_candidates", "refine_rescored", "refine_sessions", "refine_calls")}))
        ph = [int(pipe.index._lib.sss_index_stat(pipe.index._h, 9 + i)) for i in range(7)]
        if ph[0] >= 0:   # -DSSS_EXPERIMENT builds: cycles of thread 0 per refine phase, summed over invocations
            tot = float(sum(ph)) or 1.0
            print("   refine phases (init+counts, compaction, hash, k-th session, survivors, re-score, gather+sort): "
                  + " ".join("%.0f%%" % (100 * x / tot) for x in ph) + "  total %.1f Mcycles" % (tot / 1e6))
    pipe.index.set_profiling(False)


if __name__ == "__main__":
    main()
