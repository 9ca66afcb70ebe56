"""BASELINE config 5 in one script: sessions -> native featuriser -> GNN session encoder (768 -> 3 x 800 -> 3168 ->
1600) -> cosine index over the 1600-wide subsession embeddings (session max) -> top-k sessions.

    python scripts/e2e_pipeline.py [n_db_sessions] [n_query_sessions]

Database = every prefix (subsession) of every database session, contiguous per session; queries = prefixes of a sample of
database sessions, so the own session must come back first (sanity check of the whole chain, not a quality metric).
Text features are random vectors keyed by query string / item id (the reference's text model is out of scope: its
output is cached per distinct string).  Prints one JSON line with the stage throughputs.
"""
import json
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
from sessionsimilaritysearch_b200 import featurize, sessions, synth  # noqa: E402


def encode_all(enc, flat, cache, batch=200):
    out = torch.empty((len(flat), enc.out_dim), dtype=torch.float32, device=cache.device)
    t_feat = 0.0
    for lo in range(0, len(flat), batch):
        hi = min(len(flat), lo + batch)
        t0 = time.perf_counter()
        b = featurize.featurize_batch(flat.slice(lo, hi), cache)
        t_feat += time.perf_counter() - t0
        out[lo:hi] = enc(b)
    torch.cuda.synchronize()
    return out, t_feat


def main():
    n_db = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    dev = 0
    torch.cuda.set_device(dev)
    in_dim, hidden, n_layers, out_dim, msl = 768, 800, 3, 1600, 20
    enc = sss.SessionEncoder(ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 11), in_dim=in_dim, hidden=hidden,
                             n_layers=n_layers, out_dim=out_dim, max_seq_len=msl, device=dev)
    db_sessions = synth.make_sessions(n_db, 17)

    # ---- flatten: database subsessions (every prefix, contiguous per session) and the query prefixes
    t0 = time.perf_counter()
    vocab = featurize.QueryVocab()
    subs, seg = [], [0]
    for s in db_sessions:
        subs.extend(s[:j] for j in range(1, len(s) + 1))
        seg.append(len(subs))
    seg = np.asarray(seg, dtype=np.int64)
    flat_db = featurize.flatten(subs, vocab)
    rng = np.random.default_rng(5)
    pick = rng.choice(n_db, size=n_q, replace=False)
    queries = [db_sessions[i][:max(2, (2 * len(db_sessions[i])) // 3)] for i in pick]
    flat_q = featurize.flatten(queries, vocab)
    t_flatten = time.perf_counter() - t0

    # ---- text-feature cache: one row per distinct query string / item id
    item_ids = np.unique(np.concatenate([flat_db.uniq_items, [0]]))
    g = torch.Generator().manual_seed(3)
    cache = featurize.FeatureCache(torch.randn((len(vocab), in_dim), generator=g), item_ids,
                                   torch.randn((len(item_ids), in_dim), generator=g), dev)

    # ---- encode + build the index
    enc(featurize.featurize_batch(flat_db.slice(0, 200), cache))  # warm-up (cuBLAS handles, arena)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    emb_db, t_feat_db = encode_all(enc, flat_db, cache)
    t_enc_db = time.perf_counter() - t0
    t0 = time.perf_counter()
    index = sss.build_index(emb_db, 'cos')
    index.set_segments(seg, "max")
    torch.cuda.synchronize()
    t_index = time.perf_counter() - t0

    # ---- query path, timed end to end: featurise -> encode -> normalise -> search
    index.search(sss.normalize(emb_db[:256].clone()), 100)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    emb_q, t_feat_q = encode_all(enc, flat_q, cache)
    t_enc_q = time.perf_counter() - t0
    t1 = time.perf_counter()
    D, I = index.search(sss.normalize(emb_q), 100)
    torch.cuda.synchronize()
    t_search = time.perf_counter() - t1
    t_total = time.perf_counter() - t0
    own_first = float((I[:, 0].cpu().numpy() == pick).mean())
    own_in_10 = float((I[:, :10].cpu().numpy() == pick[:, None]).any(1).mean())
    print(json.dumps({
        "config": "BASELINE configs[4] scaled: %d database sessions = %d subsession rows x %d, %d query sessions, top-100 "
                  "sessions, cosine, session max" % (n_db, len(subs), out_dim, n_q),
        "flatten_sessions_per_s": (len(subs) + n_q) / t_flatten,
        "db_encode_subsessions_per_s": len(subs) / t_enc_db,
        "db_featurize_share_of_encode": t_feat_db / t_enc_db,
        "index_build_s": t_index,
        "query_sessions_per_s_end_to_end": n_q / t_total,
        "query_encode_sessions_per_s": n_q / t_enc_q,
        "query_featurize_share_of_encode": t_feat_q / t_enc_q,
        "search_queries_per_s": n_q / t_search,
        "search_scan_variant": index.stats()["scan_variant"],
        "own_session_first": own_first, "own_session_in_top10": own_in_10}))


if __name__ == "__main__":
    main()
