"""BASELINE configs[4] as one call chain: action-tuple sessions -> native featuriser -> GNN session encoder -> cosine
index over the 1600-wide (sub)session embeddings with a per-session max -> top-k sessions, on one GPU or row-sharded
over the ranks of a torch.distributed group.

What the reference does for this (test_amazon_filterd.py:485-488,546-578): `sequence_to_graph` per session, a PyG
DataLoader of 200, `encoder(batch)`, `normalize`, `build_index`, `index.search`.  Here the encoder is data-parallel
(replicated weights, no exchange) — every rank featurises and encodes its slice of the database sessions
and keeps the result as ITS row shard of the index, so the database embeddings never travel; the query sessions are
split the same way, their embeddings are all-gathered (nq x 1600 floats), and the search is the row-sharded search of
dist.ShardedIndex (one all-gather of packed candidates + merge).
"""
import time

import numpy as np
import torch

from . import featurize
from .dist import ShardedIndex
from .index import IndexFlatIP, NORM_UTIL, normalize


def encode_sessions(enc, flat, cache, batch=200, prefetch=2):
    """FlatSessions -> [n, out_dim] embeddings on the encoder's device, batches of `batch` sessions (the reference's
    DataLoader batch size, test_amazon_filterd.py:488).  Returns (embeddings, seconds spent in the host featuriser).
    A worker thread featurises up to `prefetch` batches ahead (the native call releases the GIL) while this thread
    enqueues the encoder; nothing synchronises with the device until the NaN flags are read once at the end."""
    from concurrent.futures import ThreadPoolExecutor
    dev = torch.device("cuda", enc.device)
    out = torch.empty((len(flat), enc.out_dim), dtype=torch.float32, device=dev)
    t_feat = [0.0]
    bounds = [(lo, min(len(flat), lo + batch)) for lo in range(0, len(flat), batch)]

    def make(lo, hi):
        t0 = time.perf_counter()
        with torch.cuda.device(dev):
            b = featurize.featurize_batch(flat.slice(lo, hi), cache)
        t_feat[0] += time.perf_counter() - t0
        return b

    with ThreadPoolExecutor(max_workers=1) as pool:
        pending = [pool.submit(make, lo, hi) for lo, hi in bounds[:prefetch]]
        for i, (lo, hi) in enumerate(bounds):
            b = pending.pop(0).result()
            if i + prefetch < len(bounds):
                pending.append(pool.submit(make, *bounds[i + prefetch]))
            out[lo:hi] = enc(b, defer_check=True)   # no host sync per batch
    enc.check_flags()
    return out, t_feat[0]


def rank_slice(n, rank, world):
    """contiguous, balanced [lo, hi) of n items for this rank"""
    return n * rank // world, n * (rank + 1) // world


def subsessions(sessions):
    """every prefix of every session, contiguous per session (decompose_data-style splitting): (list, seg_off)"""
    subs, seg = [], [0]
    for s in sessions:
        subs.extend(s[:j] for j in range(1, len(s) + 1))
        seg.append(len(subs))
    return subs, np.asarray(seg, dtype=np.int64)


class SessionSearchPipeline:
    """Database of sessions (rows = all their subsessions, reduced per session with max) built and searched by
    `world` ranks.  With world == 1 no process group is needed."""

    def __init__(self, enc, cache, vocab, rank=0, world=1, group=None, mode="exact"):
        self.enc, self.cache, self.vocab = enc, cache, vocab
        self.rank, self.world, self.group, self.mode = rank, world, group, mode
        self.index = None
        self.timings = {}

    def build(self, db_sessions):
        """every rank encodes the subsessions of ITS contiguous slice of the database sessions into its row shard"""
        lo, hi = rank_slice(len(db_sessions), self.rank, self.world)
        t0 = time.perf_counter()
        flat, seg = featurize.flatten_prefixes(db_sessions[lo:hi], self.vocab)
        t1 = time.perf_counter()
        emb, t_feat = encode_sessions(self.enc, flat, self.cache)
        torch.cuda.synchronize(self.enc.device)
        t2 = time.perf_counter()
        inner = IndexFlatIP(self.enc.out_dim, device=self.enc.device, id_offset=lo, mode=self.mode)
        inner.add(emb, norm=NORM_UTIL)
        inner.set_segments(seg, "max")
        torch.cuda.synchronize(self.enc.device)
        t3 = time.perf_counter()
        self.index = ShardedIndex(inner, world_size=self.world, rank=self.rank, group=self.group) if self.world > 1 else inner
        self.n_rows_local = int(seg[-1])
        self.timings.update(db_flatten_s=t1 - t0, db_encode_s=t2 - t1, db_featurize_s=t_feat, db_index_s=t3 - t2,
                            db_rows_local=self.n_rows_local, db_sessions_local=hi - lo)
        return self

    def encode_queries(self, query_sessions):
        """data-parallel encode of the query sessions; every rank ends up with all embeddings [nq, out_dim]"""
        import torch.distributed as dist
        lo, hi = rank_slice(len(query_sessions), self.rank, self.world)
        flat = featurize.flatten(query_sessions[lo:hi], self.vocab)
        emb, t_feat = encode_sessions(self.enc, flat, self.cache)
        self.timings["q_featurize_s"] = t_feat
        if self.world == 1:
            return emb
        # ragged all-gather: slices differ by at most one session
        counts = [rank_slice(len(query_sessions), r, self.world) for r in range(self.world)]
        width = max(h - l for l, h in counts)
        mine = torch.zeros((width, self.enc.out_dim), dtype=torch.float32, device=emb.device)
        mine[:emb.shape[0]] = emb
        every = torch.empty((self.world * width, self.enc.out_dim), dtype=torch.float32, device=emb.device)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        every = every.view(self.world, width, self.enc.out_dim)
        return torch.cat([every[r, :h - l] for r, (l, h) in enumerate(counts)], 0)

    def search(self, query_sessions, k=100, batch=2048):
        """query sessions -> (D [nq, k], I [nq, k]) session ids of the whole database (device tensors)"""
        emb = normalize(self.encode_queries(query_sessions))
        Ds, Is = [], []
        for lo in range(0, emb.shape[0], batch):
            D, I = self.index.search(emb[lo:lo + batch].contiguous(), k)
            Ds.append(D)
            Is.append(I)
        return torch.cat(Ds, 0), torch.cat(Is, 0)
