"""Batched featuriser: many sessions -> ONE encoder-ready batch, through the native sss_featurize_batch (C ABI).

The reference builds one PyG HeteroData per session in Python (util_amazon_filtered.py:98-230: four or five HF
tokenizer calls and a dozen small tensors each) and glues 200 of them with Batch.from_data_list
(test_amazon_filterd.py:485-488); `sessions.sequence_to_graph` + `graph.collate` mirror that call for call (about
2.2 ms per session on this image's host).  This module produces the same batch — the arrays the encoder reads,
identical node order, positions, counts, edges — from flat integer arrays in native code, and takes the node TEXT
from a feature cache keyed by query string / item id instead of tokenising it again (the text model is a pure function
of the string: SURVEY.md 8a3, 8f rank 3).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check
from .graph import EDGE_PP, EDGE_PQ, EDGE_QP, SessionBatch
from .sessions import SEARCH


class FlatSessions:
    """sessions as flat numpy arrays (the input of the native featuriser)"""

    def __init__(self, act_off, act_is_search, act_key, uniq_off, uniq_items):
        self.act_off = np.ascontiguousarray(act_off, dtype=np.int64)
        self.act_is_search = np.ascontiguousarray(act_is_search, dtype=np.uint8)
        self.act_key = np.ascontiguousarray(act_key, dtype=np.int64)
        self.uniq_off = np.ascontiguousarray(uniq_off, dtype=np.int64)
        self.uniq_items = np.ascontiguousarray(uniq_items, dtype=np.int64)

    def __len__(self):
        return len(self.act_off) - 1

    def slice(self, lo, hi):
        a0, a1 = int(self.act_off[lo]), int(self.act_off[hi])
        u0, u1 = int(self.uniq_off[lo]), int(self.uniq_off[hi])
        return FlatSessions(self.act_off[lo:hi + 1] - a0, self.act_is_search[a0:a1], self.act_key[a0:a1],
                            self.uniq_off[lo:hi + 1] - u0, self.uniq_items[u0:u1])


class QueryVocab:
    """query string -> key (row of its text feature); key 0 is the empty string of every session's root node"""

    def __init__(self):
        self.ids = {"": 0}

    def __call__(self, s):
        s = "" if s is None else s
        k = self.ids.get(s)
        if k is None:
            k = self.ids[s] = len(self.ids)
        return k

    def __len__(self):
        return len(self.ids)


def flatten(sessions, vocab, ignore_query=False):
    """action tuples (ts, type, keyword, asin, ptype, brand, title, item_id) -> FlatSessions.
    The distinct items of a session are taken as list(set(ids)), the node order of the reference
    (util_amazon_filtered.py:128)."""
    act_off, kinds, keys, uniq_off, uniq = [0], [], [], [0], []
    for seq in sessions:
        items = []
        for act in seq:
            if act[1] == SEARCH:
                if ignore_query:
                    continue
                kinds.append(1)
                keys.append(vocab(act[2]))
            else:
                kinds.append(0)
                keys.append(act[-1])
                items.append(act[-1])
        act_off.append(len(kinds))
        uniq.extend(set(items))
        uniq_off.append(len(uniq))
    return FlatSessions(act_off, kinds, keys, uniq_off, uniq)


def flatten_prefixes(sessions, vocab):
    """FlatSessions of EVERY PREFIX of every session, contiguous per session (the database rows of the session-max
    search: decompose_data-style splitting), plus seg_off [n_sessions + 1] — equal to
    flatten([s[:j] for s in sessions for j in 1..len(s)]) without materialising the prefix lists: the sessions are
    flattened once, the actions of the prefixes are gathered with index arithmetic, and only the node order of the
    distinct items (list(set(ids)) per prefix, the reference's order) is taken from Python sets."""
    base = flatten(sessions, vocab)
    n_act = np.diff(base.act_off)                              # actions per session
    seg_off = np.concatenate([[0], np.cumsum(n_act)]).astype(np.int64)   # prefixes per session = its actions
    n_pref = int(seg_off[-1])
    pref_sess = np.repeat(np.arange(len(n_act), dtype=np.int64), n_act)
    pref_len = np.arange(n_pref, dtype=np.int64) - seg_off[pref_sess] + 1          # 1 .. n per session
    act_off = np.concatenate([[0], np.cumsum(pref_len)]).astype(np.int64)
    elem_pref = np.repeat(np.arange(n_pref, dtype=np.int64), pref_len)
    src = np.arange(int(act_off[-1]), dtype=np.int64) - act_off[elem_pref] + base.act_off[pref_sess[elem_pref]]
    kinds, keys = base.act_is_search[src], base.act_key[src]
    # distinct items per prefix: consecutive prefixes between two clicks share their set
    uniq_off, uniq = [0], []
    is_search, key = base.act_is_search.tolist(), base.act_key.tolist()
    a_off = base.act_off.tolist()
    for i in range(len(n_act)):
        items, cur = [], []
        for a in range(a_off[i], a_off[i + 1]):
            if not is_search[a]:
                items.append(key[a])
                cur = list(set(items))
            uniq.extend(cur)
            uniq_off.append(len(uniq))
    return FlatSessions(act_off, kinds, keys, uniq_off, uniq), seg_off


def featurize_arrays(flat, root_query_key=0, n_threads=0):
    """native call: FlatSessions -> dict of numpy arrays (batch-global indices)"""
    lib = _lib.load()
    fs = _lib.FlatSessions(len(flat), flat.act_off.ctypes.data, flat.act_is_search.ctypes.data, flat.act_key.ctypes.data,
                           flat.uniq_off.ctypes.data, flat.uniq_items.ctypes.data)
    n = [ctypes.c_int64() for _ in range(5)]
    check(lib.sss_featurize_sizes(ctypes.byref(fs), *[ctypes.byref(x) for x in n]))
    nq, npr, ne, eqp, epp = (int(x.value) for x in n)
    a = {"query_key": np.empty(nq, np.int64), "query_pos": np.empty(nq, np.int64), "query_batch": np.empty(nq, np.int64),
         "product_key": np.empty(npr, np.int64), "product_cnt": np.empty(npr, np.int64),
         "product_batch": np.empty(npr, np.int64), "product_pos": np.empty(ne, np.int64),
         "qp_src": np.empty(eqp, np.int64), "qp_dst": np.empty(eqp, np.int64),
         "pp_src": np.empty(epp, np.int64), "pp_dst": np.empty(epp, np.int64),
         "pp_weight": np.empty(epp, np.float32), "last_click_mask": np.empty(npr, np.float32)}
    ga = _lib.GraphArrays(nq, npr, ne, eqp, epp, 0, 0, 0, 0, 0,
                          *[a[k].ctypes.data for k in ("query_key", "query_pos", "query_batch", "product_key",
                                                       "product_cnt", "product_batch", "product_pos", "qp_src",
                                                       "qp_dst", "pp_src", "pp_dst", "pp_weight", "last_click_mask")])
    check(lib.sss_featurize_batch(ctypes.byref(fs), int(root_query_key), ctypes.byref(ga), int(n_threads)))
    a["n_graphs"] = len(flat)
    return a


def featurize_group(flat, cache, batch=200, root_query_key=0, n_threads=0):
    """FlatSessions -> list of SessionBatch, one per `batch` consecutive sessions, from ONE native call
    (sss_featurize_batches), ONE host -> device copy per dtype and one feature gather per node type for the whole
    group; each returned batch is a set of views into the group's device arrays.  Equal, batch for batch, to
    [featurize_batch(flat.slice(lo, lo + batch), cache) for lo in range(0, len(flat), batch)]."""
    lib = _lib.load()
    n = len(flat)
    nb = (n + batch - 1) // batch
    if nb == 0:
        return []
    fs = _lib.FlatSessions(n, flat.act_off.ctypes.data, flat.act_is_search.ctypes.data, flat.act_key.ctypes.data,
                           flat.uniq_off.ctypes.data, flat.uniq_items.ctypes.data)
    # upper bounds instead of a sizing pass (root + searches; distinct items or one placeholder; item events or one
    # placeholder; one q->p edge per item event; fewer transitions than item events): the exact sizes come back from
    # the call and only that much is copied to the device
    n_act, n_uniq = int(flat.act_off[-1]), int(flat.uniq_off[-1])
    nq, npr, ne, eqp, epp = n + n_act, n_uniq + n, n_act + n, n_act, n_act
    # one int64 slab: [query_key | query_pos | query_batch | product_key | product_cnt | product_batch | product_pos |
    #                  item_rows | qp_src | qp_dst | pp_src | pp_dst]; one fp32 slab: [pp_weight | last_click_mask]
    names = ("query_key", "query_pos", "query_batch", "product_key", "product_cnt", "product_batch", "product_pos",
             "item_rows", "qp_src", "qp_dst", "pp_src", "pp_dst")
    sizes = (nq, nq, nq, npr, npr, npr, ne, npr, eqp, eqp, epp, epp)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    slab = np.empty(int(starts[-1]), np.int64)
    a = {k: slab[starts[i]:starts[i + 1]] for i, k in enumerate(names)}
    cap_epp = epp
    fslab = np.empty(epp + npr, np.float32)
    a["pp_weight"], a["last_click_mask"] = fslab[:epp], fslab[epp:]
    ga = _lib.GraphArrays(nq, npr, ne, eqp, epp, 0, 0, 0, 0, 0,
                          *[a[k].ctypes.data for k in ("query_key", "query_pos", "query_batch", "product_key",
                                                       "product_cnt", "product_batch", "product_pos", "qp_src",
                                                       "qp_dst", "pp_src", "pp_dst", "pp_weight", "last_click_mask")])
    bounds = np.empty((nb + 1, 5), np.int64)
    check(lib.sss_featurize_batches(ctypes.byref(fs), int(batch), int(root_query_key), ctypes.byref(ga),
                                    bounds.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), int(n_threads)))
    nq, npr, ne, eqp, epp = (int(v) for v in bounds[nb])          # exact sizes
    real = dict(zip(names, (nq, nq, nq, npr, npr, npr, ne, npr, eqp, eqp, epp, epp)))
    a = {k: (v[:real[k]] if k in real else v) for k, v in a.items()}
    a["item_rows"][:] = cache.item_rows(a["product_key"])
    dev = cache.device
    d_i = torch.from_numpy(slab).to(dev, non_blocking=True)
    d_f = torch.from_numpy(fslab).to(dev, non_blocking=True)
    dv = {k: d_i[starts[i]:starts[i] + real[k]] for i, k in enumerate(names)}
    dv["pp_weight"], dv["last_click_mask"] = d_f[:epp], d_f[cap_epp:cap_epp + npr]
    if cache.query_features.device.type == "cuda":
        from .encoder import gather_rows
        xq_all = gather_rows(cache.query_features, dv["query_key"])
        xp_all = gather_rows(cache.item_features, dv["item_rows"])
    else:
        xq_all = cache.query_features.index_select(0, dv["query_key"])
        xp_all = cache.item_features.index_select(0, dv["item_rows"])
    qp_all = torch.stack([dv["qp_src"], dv["qp_dst"]])
    pq_all = torch.stack([dv["qp_dst"], dv["qp_src"]])
    pp_all = torch.stack([dv["pp_src"], dv["pp_dst"]])
    out = []
    bl = bounds.tolist()
    for b in range(nb):
        (q0, p0, e0, g0, h0), (q1, p1, e1, g1, h1) = bl[b], bl[b + 1]
        sb = SessionBatch()
        sb.num_graphs = min(batch, n - b * batch)
        q = sb["query"]
        q.x = xq_all[q0:q1]
        q.pos_emb_id = dv["query_pos"][q0:q1]
        q.batch = dv["query_batch"][q0:q1]
        q.num_nodes = q1 - q0
        p = sb["product"]
        p.x = dv["product_key"][p0:p1]
        p.input_ids = xp_all[p0:p1]
        p.cnt = dv["product_cnt"][p0:p1]
        p.pos_emb_id = dv["product_pos"][e0:e1]
        p.batch = dv["product_batch"][p0:p1]
        p.last_click_mask = dv["last_click_mask"][p0:p1]
        p.num_nodes = p1 - p0
        sb[EDGE_QP].edge_index = qp_all[:, g0:g1]
        sb[EDGE_PQ].edge_index = pq_all[:, g0:g1]
        pp = sb[EDGE_PP]
        pp.edge_index = pp_all[:, h0:h1]
        pp.edge_weight = dv["pp_weight"][h0:h1]
        out.append(sb)
    return out


class FeatureCache:
    """text features on the device, one row per query key and one per item id (what the reference's
    PretrainedQAEAEncoder computes per node, model/NodeEmbedding.py:112-125, computed once per distinct string)"""

    def __init__(self, query_features, item_ids, item_features, device):
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.device = dev
        self.query_features = torch.as_tensor(query_features, dtype=torch.float32).to(dev)
        ids = np.asarray(item_ids, dtype=np.int64)
        self.item_features = torch.as_tensor(item_features, dtype=torch.float32).to(dev)
        order = np.argsort(ids, kind="stable")
        self._sorted_ids = ids[order]
        self._rows = order

    def item_rows(self, item_ids):
        pos = np.searchsorted(self._sorted_ids, item_ids)
        pos = np.minimum(pos, len(self._sorted_ids) - 1)
        if not np.array_equal(self._sorted_ids[pos], item_ids):
            raise KeyError("item id without a cached text feature")
        return self._rows[pos]


def featurize_batch(flat, cache, root_query_key=0, n_threads=0):
    """FlatSessions -> SessionBatch on the cache's device, ready for SessionEncoder.__call__ (same attributes as
    graph.collate([sequence_to_graph(...), ...]) with the text features already in place)."""
    a = featurize_arrays(flat, root_query_key, n_threads)
    dev = cache.device
    # ONE host -> device copy per dtype for the whole batch (a dozen small pageable copies cost more than the native
    # featuriser itself): the arrays are laid end to end and sliced on the device
    i_keys = ("query_key", "query_pos", "query_batch", "product_key", "product_cnt", "product_batch", "product_pos",
              "qp_src", "qp_dst", "pp_src", "pp_dst")
    a["item_rows"] = cache.item_rows(a["product_key"]).astype(np.int64, copy=False)
    i_keys = i_keys + ("item_rows",)
    f_keys = ("pp_weight", "last_click_mask")
    i_all = torch.from_numpy(np.concatenate([a[k] for k in i_keys])).to(dev, non_blocking=True)
    f_all = torch.from_numpy(np.concatenate([a[k] for k in f_keys])).to(dev, non_blocking=True)
    dv, off = {}, 0
    for k in i_keys:
        dv[k] = i_all[off:off + len(a[k])]
        off += len(a[k])
    off = 0
    for k in f_keys:
        dv[k] = f_all[off:off + len(a[k])]
        off += len(a[k])

    def t(x):
        return x

    def rows(table, idx):  # feature gather: the library's kernel on the device, plain indexing for a host cache
        if table.device.type == "cuda":
            from .encoder import gather_rows
            return gather_rows(table, idx)
        return table.index_select(0, idx)

    out = SessionBatch()
    out.num_graphs = a["n_graphs"]
    q = out["query"]
    q.x = rows(cache.query_features, dv["query_key"])
    q.pos_emb_id = dv["query_pos"]
    q.batch = dv["query_batch"]
    q.num_nodes = len(a["query_key"])
    p = out["product"]
    p.x = dv["product_key"]
    p.input_ids = rows(cache.item_features, dv["item_rows"])
    p.cnt = dv["product_cnt"]
    p.pos_emb_id = dv["product_pos"]
    p.batch = dv["product_batch"]
    p.last_click_mask = dv["last_click_mask"]
    p.num_nodes = len(a["product_key"])
    qp = torch.stack([dv["qp_src"], dv["qp_dst"]])
    out[EDGE_QP].edge_index = qp
    out[EDGE_PQ].edge_index = qp.flip(0)
    pp = out[EDGE_PP]
    pp.edge_index = torch.stack([dv["pp_src"], dv["pp_dst"]])
    pp.edge_weight = dv["pp_weight"]
    return out
