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
import contextlib
import gc
import time

import numpy as np
import torch

from . import featurize
from .dist import ShardedIndex
from .index import IndexFlatIP, NORM_UTIL, normalize


def encode_sessions(enc, flat, cache, batch=200, group=20):
    """FlatSessions -> [n, out_dim] embeddings on the encoder's device, batches of `batch` sessions (the reference's
    DataLoader batch size, test_amazon_filterd.py:488).  Returns (embeddings, seconds spent in the host featuriser).
    A worker thread featurises `group` batches per native call, one group ahead of the encoder (the native call and the
    copies release the GIL); nothing synchronises with the device until the NaN flags are read once at the end."""
    from concurrent.futures import ThreadPoolExecutor
    dev = torch.device("cuda", enc.device)
    out = torch.empty((len(flat), enc.out_dim), dtype=torch.float32, device=dev)
    t_feat = [0.0]
    step = batch * group
    bounds = [(lo, min(len(flat), lo + step)) for lo in range(0, len(flat), step)]

    def make(lo, hi):
        t0 = time.perf_counter()
        with torch.cuda.device(dev):
            bs = featurize.featurize_group(flat.slice(lo, hi), cache, batch)
        t_feat[0] += time.perf_counter() - t0
        return bs

    with ThreadPoolExecutor(max_workers=1) as pool:
        nxt = pool.submit(make, *bounds[0]) if bounds else None
        for i, (lo, hi) in enumerate(bounds):
            batches = nxt.result()
            nxt = pool.submit(make, *bounds[i + 1]) if i + 1 < len(bounds) else None
            for j, b in enumerate(batches):
                r0 = lo + j * batch
                out[r0:min(hi, r0 + batch)] = enc(b, defer_check=True)   # no host sync per batch
    enc.check_flags()
    return out, t_feat[0]


def encode_session_lists(enc, sessions, vocab, cache, prefixes=False, batch=200, group=20):
    """action-tuple sessions -> embeddings, with the WHOLE host side (flattening included) on the worker thread:
    chunk i + 1 is flattened and featurised while the GPU encodes chunk i.  prefixes=True encodes every prefix of every
    session (rows contiguous per session) and also returns seg_off.  Returns (embeddings, seg_off or None, host seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    dev = torch.device("cuda", enc.device)
    lens = np.fromiter((len(s) for s in sessions), dtype=np.int64, count=len(sessions))
    rows_of = lens if prefixes else np.ones(len(sessions), dtype=np.int64)
    row_off = np.concatenate([[0], np.cumsum(rows_of)]).astype(np.int64)
    out = torch.empty((int(row_off[-1]), enc.out_dim), dtype=torch.float32, device=dev)
    # chunks of exactly batch * group ROWS (the batches are then the reference's DataLoader batches: rows [0, 200),
    # [200, 400), ... — GATConv's bipartite self-loop quirk makes an embedding depend on its batch);
    # a session whose prefixes straddle a chunk boundary is flattened for both chunks
    n_rows, target = int(row_off[-1]), batch * group
    cuts = list(range(0, n_rows, target)) + [n_rows]
    t_host = [0.0]

    def make(r_lo, r_hi):
        t0 = time.perf_counter()
        s_lo = int(np.searchsorted(row_off, r_lo, side="right")) - 1
        s_hi = int(np.searchsorted(row_off, r_hi, side="left"))
        if prefixes:
            flat, _ = featurize.flatten_prefixes(sessions[s_lo:s_hi], vocab)
        else:
            flat = featurize.flatten(sessions[s_lo:s_hi], vocab)
        first = r_lo - int(row_off[s_lo])
        flat = flat.slice(first, first + (r_hi - r_lo))
        with torch.cuda.device(dev):
            bs = featurize.featurize_group(flat, cache, batch)
        t_host[0] += time.perf_counter() - t0
        return bs

    with ThreadPoolExecutor(max_workers=1) as pool:
        nxt = pool.submit(make, cuts[0], cuts[1]) if len(cuts) > 1 else None
        for i in range(len(cuts) - 1):
            batches = nxt.result()
            nxt = pool.submit(make, cuts[i + 1], cuts[i + 2]) if i + 2 < len(cuts) else None
            r0 = cuts[i]
            for b in batches:
                out[r0:r0 + b.num_graphs] = enc(b, defer_check=True)   # no host sync per batch
                r0 += b.num_graphs
    enc.check_flags()
    return out, (row_off if prefixes else None), t_host[0]


@contextlib.contextmanager
def _gc_paused():
    """The chain's hot loops allocate many small Python objects next to the caller's millions of action tuples; a full
    collection of the cyclic GC in the middle of them walks all of those (tens of ms per search batch, measured).
    Paused for the duration of a call, restored afterwards."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


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
        with _gc_paused():
            return self._build(db_sessions)

    def _build(self, db_sessions):
        lo, hi = rank_slice(len(db_sessions), self.rank, self.world)
        t0 = t1 = time.perf_counter()
        emb, seg, t_feat = encode_session_lists(self.enc, db_sessions[lo:hi], self.vocab, self.cache, prefixes=True)
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
        emb, _, t_feat = encode_session_lists(self.enc, query_sessions[lo:hi], self.vocab, self.cache)
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
        with _gc_paused():
            return self._search(query_sessions, k, batch)

    def _search(self, query_sessions, k, batch):
        t0 = time.perf_counter()
        emb = normalize(self.encode_queries(query_sessions))
        torch.cuda.synchronize(self.enc.device)
        t1 = time.perf_counter()
        # results are written straight into the final tensors: no allocation inside the loop
        D = torch.empty((emb.shape[0], k), dtype=torch.float32, device=emb.device)
        I = torch.empty((emb.shape[0], k), dtype=torch.int64, device=emb.device)
        per_batch = []
        for lo in range(0, emb.shape[0], batch):
            tb = time.perf_counter()
            hi = min(emb.shape[0], lo + batch)
            self.index.search(emb[lo:hi], k, out=(D[lo:hi], I[lo:hi]))
            per_batch.append((time.perf_counter() - tb) * 1e3)   # (a search call returns after its status read-back)
        torch.cuda.synchronize(self.enc.device)
        self.timings.update(q_encode_s=t1 - t0, q_search_s=time.perf_counter() - t1, q_search_batch_ms=per_batch)
        return D, I
