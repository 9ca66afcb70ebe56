"""Native batched featuriser (sss_featurize_batch, host code of libsss_b200.so) against
(1) the graphs the REFERENCE's own sequence_to_graph built (tests/golden/graphs_golden.npz) and
(2) the Python mirror + collate on larger seeded batches, array for array."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from sessionsimilaritysearch_b200 import featurize, graph, sessions, synth  # noqa: E402

GG = np.load(os.path.join(HERE, "golden", "graphs_golden.npz"))


def _sessions_of_the_goldens():
    sess = synth.make_sessions(24, 5)
    sess[3] = [a for a in sess[3] if a[1] == 's'] or sess[3]
    return sess


@pytest.mark.parametrize("ignore_query", [False, True])
def test_native_featuriser_matches_the_reference_graphs(ignore_query):
    sess = _sessions_of_the_goldens()
    if ignore_query:
        sess = sess[:6]
    prefix = "ig_g%d" if ignore_query else "g%d"
    flat = featurize.flatten(sess, featurize.QueryVocab(), ignore_query=ignore_query)
    a = featurize.featurize_arrays(flat)
    q0 = p0 = e0 = qp0 = pp0 = 0
    for i in range(len(sess)):
        g = prefix % i
        nq = len(GG[g + "_query_pos_emb_id"])
        npn = len(GG[g + "_product_x"])
        ne = len(GG[g + "_product_pos_emb_id"])
        assert np.array_equal(a["query_pos"][q0:q0 + nq], GG[g + "_query_pos_emb_id"])
        assert np.all(a["query_batch"][q0:q0 + nq] == i)
        assert np.array_equal(a["product_key"][p0:p0 + npn], GG[g + "_product_x"])
        assert np.array_equal(a["product_cnt"][p0:p0 + npn], GG[g + "_product_cnt"])
        assert np.array_equal(a["last_click_mask"][p0:p0 + npn], GG[g + "_product_last_click_mask"])
        assert np.array_equal(a["product_pos"][e0:e0 + ne], GG[g + "_product_pos_emb_id"])
        qp = GG[g + "_qp"]
        n_qp = qp.shape[1]
        assert np.array_equal(a["qp_src"][qp0:qp0 + n_qp] - q0, qp[0])
        assert np.array_equal(a["qp_dst"][qp0:qp0 + n_qp] - p0, qp[1])
        pp = GG[g + "_pp"]
        n_pp = pp.shape[1]
        assert np.array_equal(a["pp_src"][pp0:pp0 + n_pp] - p0, pp[0])
        assert np.array_equal(a["pp_dst"][pp0:pp0 + n_pp] - p0, pp[1])
        assert np.array_equal(a["pp_weight"][pp0:pp0 + n_pp], GG[g + "_pp_w"])
        q0, p0, e0, qp0, pp0 = q0 + nq, p0 + npn, e0 + ne, qp0 + n_qp, pp0 + n_pp
    assert (q0, p0, e0, qp0, pp0) == (len(a["query_pos"]), len(a["product_key"]), len(a["product_pos"]),
                                      len(a["qp_src"]), len(a["pp_src"]))


@pytest.mark.parametrize("n,threads", [(1, 1), (700, 1), (6000, 4)])
def test_native_featuriser_equals_python_mirror_plus_collate(n, threads):
    tok = synth.HashTokenizer()
    sess = synth.make_sessions(n, 100 + n)
    if n > 3:
        sess[2] = [a for a in sess[2] if a[1] == 's'] or sess[2]          # item-less
        sess[3] = [a for a in sess[3] if a[1] != 's'] or sess[3]          # search-less
    vocab = featurize.QueryVocab()
    a = featurize.featurize_arrays(featurize.flatten(sess, vocab), n_threads=threads)
    b = graph.collate([sessions.sequence_to_graph(0, s, s[:1], tok, 20) for s in sess])
    ei = b.edge_index_dict
    want = {"query_pos": b["query"].pos_emb_id, "query_batch": b["query"].batch, "product_key": b["product"].x,
            "product_cnt": b["product"].cnt, "product_pos": b["product"].pos_emb_id,
            "product_batch": b["product"].batch, "last_click_mask": b["product"].last_click_mask,
            "qp_src": ei[graph.EDGE_QP][0], "qp_dst": ei[graph.EDGE_QP][1], "pp_src": ei[graph.EDGE_PP][0],
            "pp_dst": ei[graph.EDGE_PP][1], "pp_weight": b[graph.EDGE_PP].edge_weight}
    for k, v in want.items():
        assert np.array_equal(a[k], v.numpy()), k
    assert np.array_equal(np.stack([a["qp_dst"], a["qp_src"]]), ei[graph.EDGE_PQ].numpy())
    # the text key of a query node is the vocabulary id of its string; node 0 of every session is the empty string
    words = [w for s in sess
             for w in [""] + [(act[2] if act[2] is not None else "") for act in s if act[1] == sessions.SEARCH]]
    assert [vocab(w) for w in words] == a["query_key"].tolist()


def test_native_featuriser_rejects_inconsistent_input():
    flat = featurize.flatten(synth.make_sessions(5, 1), featurize.QueryVocab())
    flat.uniq_items[0] += 12345  # an item event whose id is not in the session's distinct list
    with pytest.raises(RuntimeError, match="distinct-item list"):
        featurize.featurize_arrays(flat)


def test_flat_sessions_slice():
    sess = synth.make_sessions(50, 9)
    flat = featurize.flatten(sess, featurize.QueryVocab())
    part = featurize.featurize_arrays(flat.slice(10, 30))
    whole = featurize.featurize_arrays(featurize.flatten(sess[10:30], featurize.QueryVocab()))
    for k in ("query_pos", "product_key", "product_pos", "qp_src", "pp_dst"):
        assert np.array_equal(part[k], whole[k])


def test_featurize_batch_with_a_feature_cache_matches_collate_on_the_host():
    """feature gather through the cache (device = cpu here): same node features, in the same node order, as
    tokenising every node of every session and looking its feature up (the synthetic stand-in of the text model)"""
    import torch
    in_dim = 16
    tok = synth.HashTokenizer()
    sess = synth.make_sessions(60, 77)
    sess[5] = [a for a in sess[5] if a[1] == 's'] or sess[5]

    def feats(strings):
        ids = tok(strings, padding='max_length', max_length=20, truncation=True, return_tensors="pt")['input_ids']
        return synth.text_features(ids, in_dim)

    vocab = featurize.QueryVocab()
    flat = featurize.flatten(sess, vocab)
    titles = {0: 'UNK'}
    for s in sess:
        for act in s:
            if act[1] != sessions.SEARCH:
                titles.setdefault(act[-1], act[-2] if act[-2] is not None else '')
    item_ids = sorted(titles)
    cache = featurize.FeatureCache(feats(sorted(vocab.ids, key=vocab.ids.get)), item_ids,
                                   feats([titles[i] for i in item_ids]), "cpu")
    got = featurize.featurize_batch(flat, cache)
    graphs = [sessions.sequence_to_graph(0, s, s[:1], tok, 20) for s in sess]
    for g in graphs:
        g['query'].x = synth.text_features(g['query'].input_ids, in_dim)
        g['product'].input_ids = synth.text_features(g['product'].input_ids, in_dim)
    want = graph.collate(graphs)
    assert torch.equal(got['query'].x, want['query'].x)
    assert torch.equal(got['product'].input_ids, want['product'].input_ids)
    assert torch.equal(got['product'].cnt, want['product'].cnt)
    assert torch.equal(got.edge_index_dict[graph.EDGE_PQ], want.edge_index_dict[graph.EDGE_PQ])
    assert got.num_graphs == want.num_graphs == 60
    # an item without a cached feature is an error, not a silent zero row
    small = featurize.FeatureCache(feats([""] * len(vocab)), item_ids[:3], feats(["x"] * 3), "cpu")
    with pytest.raises(KeyError):
        featurize.featurize_batch(flat, small)


def test_flatten_prefixes_equals_flatten_of_the_prefix_lists():
    from sessionsimilaritysearch_b200 import pipeline
    sess = synth.make_sessions(60, 5)
    v1, v2 = featurize.QueryVocab(), featurize.QueryVocab()
    subs, seg = pipeline.subsessions(sess)
    ref = featurize.flatten(subs, v1)
    got, seg2 = featurize.flatten_prefixes(sess, v2)
    assert np.array_equal(seg, seg2) and v1.ids == v2.ids
    for name in ("act_off", "act_is_search", "act_key", "uniq_off", "uniq_items"):
        assert np.array_equal(getattr(ref, name), getattr(got, name)), name


def test_featurize_group_equals_one_featurize_batch_per_slice():
    """sss_featurize_batches (many encoder batches per native call, batch-local indices, views into one slab) against
    one sss_featurize_batch call per batch, every attribute the encoder reads"""
    import torch
    in_dim = 8
    sess = synth.make_sessions(53, 9)            # 53 sessions in batches of 10: a ragged last batch
    sess[7] = [a for a in sess[7] if a[1] == 's'] or sess[7]     # an item-less session
    vocab = featurize.QueryVocab()
    flat = featurize.flatten(sess, vocab)
    item_ids = sorted({0} | {a[-1] for s in sess for a in s if a[1] != sessions.SEARCH})
    g = torch.Generator().manual_seed(3)
    cache = featurize.FeatureCache(torch.randn(len(vocab), in_dim, generator=g), item_ids,
                                   torch.randn(len(item_ids), in_dim, generator=g), "cpu")
    got = featurize.featurize_group(flat, cache, batch=10)
    assert len(got) == 6
    for b, lo in enumerate(range(0, 53, 10)):
        want = featurize.featurize_batch(flat.slice(lo, min(53, lo + 10)), cache)
        assert got[b].num_graphs == want.num_graphs
        for node in ("query", "product"):
            for attr in ("pos_emb_id", "batch"):
                assert torch.equal(getattr(got[b][node], attr), getattr(want[node], attr)), (b, node, attr)
        assert torch.equal(got[b]["query"].x, want["query"].x)
        assert torch.equal(got[b]["product"].x, want["product"].x)
        assert torch.equal(got[b]["product"].input_ids, want["product"].input_ids)
        assert torch.equal(got[b]["product"].cnt, want["product"].cnt)
        assert torch.equal(got[b]["product"].last_click_mask, want["product"].last_click_mask)
        for key in (graph.EDGE_QP, graph.EDGE_PQ, graph.EDGE_PP):
            assert torch.equal(got[b].edge_index_dict[key], want.edge_index_dict[key]), (b, key)
        assert torch.equal(got[b][graph.EDGE_PP].edge_weight, want[graph.EDGE_PP].edge_weight)


def test_native_featuriser_worker_pool_under_concurrent_callers():
    """the persistent worker pool of sss_featurize_batch(es): one parallel region at a time, a caller that finds it busy
    runs inline — four Python threads calling at once (the GIL is released inside the call) must all get the
    single-threaded result, call after call"""
    from concurrent.futures import ThreadPoolExecutor
    sess = synth.make_sessions(3000, 21)
    flat = featurize.flatten(sess, featurize.QueryVocab())
    want = featurize.featurize_arrays(flat, n_threads=1)

    def once(threads):
        got = featurize.featurize_arrays(flat, n_threads=threads)
        return all(np.array_equal(got[k], want[k]) for k in want if k != "n_graphs")

    with ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(once, [8, 4, 0, 2] * 6))
    assert all(results)
    assert once(0) and once(3)
