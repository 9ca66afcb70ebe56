"""GPU parity tests of the encoder path (sss_encoder_forward, sss_binarize_head, sss_item_vote) against the CPU
oracle and the goldens produced by the reference's own model code."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import encoder_common as ec  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", ec.CONFIGS, ids=[c[0] for c in ec.CONFIGS])
def test_encoder_matches_reference_golden_and_oracle(cfg):
    import sessionsimilaritysearch_b200 as sss
    from oracle import encoder_oracle as eo
    from sessionsimilaritysearch_b200 import graph, sessions
    name, in_dim, hidden, n_layers, out_dim, msl, n_sess, seed = cfg
    gold = np.load(os.path.join(HERE, "golden", "encoder_golden_%s.npz" % name))
    _, graphs = ec.make_graphs(n_sess, in_dim, seed, sessions.sequence_to_graph)
    data = graph.collate(graphs)
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, seed)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    out = enc(data.to("cuda")).cpu().numpy()
    scale = float(np.abs(gold["out"]).max())
    # split-bf16 tensor-core linears against torch's fp32: tolerance 2e-4 of the output scale
    np.testing.assert_allclose(out, gold["out"], rtol=2e-4, atol=2e-4 * scale)
    ref = eo.encoder_forward(P, eo.batch_from_pyg(graph.collate(graphs)), n_layers).numpy()
    np.testing.assert_allclose(out, ref, rtol=2e-4, atol=2e-4 * scale)
    # deterministic run to run
    assert np.array_equal(out, enc(data).cpu().numpy())
    # hash head on the golden embeddings reproduces the reference's codes exactly where |pre-activation| is not ~0
    head = sss.BinarizeHead(torch.from_numpy(gold["head_w"]), torch.from_numpy(gold["head_b"]))
    codes = head(torch.from_numpy(gold["out"])).cpu().numpy()
    pre = gold["out"] @ gold["head_w"].T + gold["head_b"]
    safe = np.abs(pre) > 1e-4 * np.abs(pre).max()
    assert np.array_equal(codes[safe], gold["codes"][safe]) and set(np.unique(codes)) <= {-1.0, 0.0, 1.0}


def test_encoder_larger_batch_against_oracle():
    import sessionsimilaritysearch_b200 as sss
    from oracle import encoder_oracle as eo
    from sessionsimilaritysearch_b200 import graph, sessions
    in_dim, hidden, n_layers, out_dim, msl = 48, 64, 3, 100, 20
    _, graphs = ec.make_graphs(200, in_dim, 3, sessions.sequence_to_graph)     # the reference's eval batch size
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 3)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    loader = graph.DataLoader(graphs, batch_size=200, shuffle=False)
    data = next(iter(loader))
    out = enc(data.to("cuda")).cpu().numpy()
    ref = eo.encoder_forward(P, eo.batch_from_pyg(graph.collate(graphs)), n_layers).numpy()
    np.testing.assert_allclose(out, ref, rtol=3e-4, atol=3e-4 * float(np.abs(ref).max()))
    assert out.shape == (200, out_dim)


def test_encoder_runs_on_its_own_gemm_only():
    """the encoder's linears run on this library's split-bf16 tcgen05 GEMM with fused epilogues — 16 launches per
    forward and no library GEMM; the cuBLAS arithmetics of earlier versions are rejected"""
    import sessionsimilaritysearch_b200 as sss
    from sessionsimilaritysearch_b200 import graph, sessions
    in_dim, hidden, n_layers, out_dim, msl = 768, 800, 3, 1600, 20
    _, graphs = ec.make_graphs(40, in_dim, 5, sessions.sequence_to_graph)
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 5)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    assert enc.math == "bf16x3"
    for m in ("fp32", "bf16x9"):
        with pytest.raises(RuntimeError):
            enc.set_math(m)
    enc(graph.collate(graphs).to("cuda"))
    assert enc.launches <= 16, enc.launches


def test_native_featuriser_feeds_the_encoder_like_the_python_path():
    """flatten -> sss_featurize_batch -> feature cache gather -> encoder must give the very same embeddings as
    sequence_to_graph per session -> collate -> encoder (same node order, same features, same kernels)."""
    import torch
    import sessionsimilaritysearch_b200 as sss
    from sessionsimilaritysearch_b200 import featurize, graph, sessions, synth
    in_dim, hidden, n_layers, out_dim, msl = 48, 64, 3, 100, 20
    sess, graphs = ec.make_graphs(300, in_dim, 23, sessions.sequence_to_graph)
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 23)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    tok = synth.HashTokenizer()

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
                                   feats([titles[i] for i in item_ids]), 0)
    out_native = enc(featurize.featurize_batch(flat, cache))
    out_python = enc(graph.collate(graphs).to("cuda"))
    assert torch.equal(out_native, out_python)
    # (a sub-batch is its own batch: the GAT self-loop quirk ties embeddings to the batch composition)
    part = enc(featurize.featurize_batch(flat.slice(100, 200), cache))
    assert torch.equal(part, enc(graph.collate(graphs[100:200]).to("cuda")))


def test_chunked_threaded_pipeline_encode_equals_batch_by_batch():
    """pipeline.encode_session_lists (host side on a worker thread, many batches per native call, chunks cut by rows)
    against one featurize_batch + forward per DataLoader batch: bit-identical, for plain sessions and for every prefix"""
    import torch
    import sessionsimilaritysearch_b200 as sss
    from sessionsimilaritysearch_b200 import featurize, pipeline, sessions, synth
    in_dim, hidden, n_layers, out_dim, msl = 48, 64, 2, 100, 20
    sess = synth.make_sessions(230, 41)
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 41)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    vocab = featurize.QueryVocab()
    featurize.flatten(sess, vocab)
    item_ids = sorted({0} | {a[-1] for s in sess for a in s if a[1] != sessions.SEARCH})
    g = torch.Generator().manual_seed(4)
    cache = featurize.FeatureCache(torch.randn(len(vocab), in_dim, generator=g), item_ids,
                                   torch.randn(len(item_ids), in_dim, generator=g), 0)
    for prefixes in (False, True):
        rows = pipeline.subsessions(sess)[0] if prefixes else sess
        flat = featurize.flatten(rows, vocab)
        want = torch.cat([enc(featurize.featurize_batch(flat.slice(lo, min(len(flat), lo + 32)), cache))
                          for lo in range(0, len(flat), 32)])
        got, seg, _ = pipeline.encode_session_lists(enc, sess, vocab, cache, prefixes=prefixes, batch=32, group=3)
        assert torch.equal(got, want), prefixes
        if prefixes:
            assert np.array_equal(seg, pipeline.subsessions(sess)[1])
        old, _ = pipeline.encode_sessions(enc, flat, cache, batch=32, group=3)
        assert torch.equal(old, want)


@pytest.mark.parametrize("shape", [(768, 800, 3, 1600, 40), (768, 800, 3, 1600, 200), (48, 64, 3, 100, 200),
                                   (24, 40, 2, 52, 30)])
def test_encoder_tcgen05_linears_against_float64(shape):
    """the fused tcgen05 forward against a FLOAT64 run of the oracle (the arbiter between fp32 summation orders):
    inside 1e-4 of the output scale (measured 2.7e-5 at the model shape), deterministic run to run; shapes whose K
    slices are not whole MMA steps (24 / 40 / 52) included"""
    import sessionsimilaritysearch_b200 as sss
    from oracle import encoder_oracle as eo
    from sessionsimilaritysearch_b200 import graph, sessions
    in_dim, hidden, n_layers, out_dim, n_sess = shape
    msl = 20
    _, graphs = ec.make_graphs(n_sess, in_dim, 5, sessions.sequence_to_graph)
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 5)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    data = graph.collate(graphs).to("cuda")
    out3 = enc(data).cpu().numpy()
    ref = eo.encoder_forward(P, eo.batch_from_pyg(graph.collate(graphs)), n_layers).numpy()
    P64 = {k: v.double() for k, v in P.items()}
    ref64 = eo.encoder_forward(P64, eo.batch_from_pyg(graph.collate(graphs)), n_layers).numpy()
    scale = float(np.abs(ref64).max())
    np.testing.assert_allclose(out3, ref, rtol=3e-4, atol=3e-4 * scale)
    assert float(np.abs(out3 - ref64).max()) <= 1e-4 * scale, float(np.abs(out3 - ref64).max()) / scale
    assert np.array_equal(enc(data).cpu().numpy(), out3)


def test_gather_rows_and_id_embedding():
    """sss_gather_rows = nn.Embedding lookup (NodeAsinEmbedding, model/NodeEmbedding.py:137-138): bit-equal to torch
    indexing, odd widths included, out-of-range ids rejected with torch's message"""
    import torch
    import sessionsimilaritysearch_b200 as sss
    g = torch.Generator().manual_seed(1)
    for n_rows, d, n in ((1000, 200, 777), (50, 7, 33), (391, 768, 5)):
        table = torch.randn((n_rows, d), generator=g)
        ids = torch.randint(0, n_rows, (n,), generator=g)
        emb = sss.NodeAsinEmbedding(table)
        assert torch.equal(emb(ids).cpu(), table[ids])
    with pytest.raises(RuntimeError, match="index out of range"):
        emb(torch.tensor([0, 391]))
    assert emb(torch.zeros(0, dtype=torch.int64)).shape == (0, 768)


def test_nan_input_raises_like_the_reference():
    import sessionsimilaritysearch_b200 as sss
    from sessionsimilaritysearch_b200 import graph, sessions
    in_dim, hidden, n_layers, out_dim, msl = 24, 32, 3, 60, 20
    _, graphs = ec.make_graphs(4, in_dim, 7, sessions.sequence_to_graph)
    graphs[2]['query'].x[0, 3] = float("nan")
    enc = sss.SessionEncoder(ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 7), in_dim=in_dim, hidden=hidden,
                             n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    with pytest.raises(RuntimeError):
        enc(graph.collate(graphs).to("cuda"))


def test_item_vote_and_get_prediction_by_knn(oracle):
    import sessionsimilaritysearch_b200 as sss
    rng = np.random.default_rng(14)
    n_sess = 400
    lens = rng.integers(1, 9, size=n_sess)
    item_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    items = rng.integers(0, 300, size=item_off[-1]).astype(np.int64)
    lists = sss.ItemLists([items[item_off[i]:item_off[i + 1]] for i in range(n_sess)])
    I = np.stack([rng.permutation(n_sess)[:60] for _ in range(9)]).astype(np.int64)
    I[0, 50:] = -1                       # padded neighbours are skipped
    D = np.sort(rng.random((9, 60)).astype(np.float32), axis=1)[:, ::-1].copy()
    # (reference semantics — float64 sums, ties in arrival order — are pinned in test_metrics_gpu.py against the
    # fixture the reference's own function produced)
    oi, ow = sss.item_vote(D, I, lists, 20)
    ei, ew = oracle.item_vote(D, I, item_off, items, 20)
    assert np.array_equal(oi.cpu().numpy(), ei)
    assert np.array_equal(ow.cpu().numpy().view(np.uint32), ew.view(np.uint32))
    # end to end like the reference: search 60 neighbours, vote, top 20 items
    emb = rng.standard_normal((n_sess, 32)).astype(np.float32)
    index = sss.build_index(emb, 'cos')

    class G(dict):
        pass
    dataset = [G(product=type("P", (), {"x": torch.from_numpy(items[item_off[i]:item_off[i + 1]])})()) for i in range(n_sess)]
    q = emb[7] + 0.01 * rng.standard_normal(32).astype(np.float32)
    pred = sss.get_prediction_by_knn(torch.from_numpy(q), index, dataset, 60, 20)
    Dq, Iq = index.search(q[None], 60)
    ei, _ = oracle.item_vote(Dq, Iq, item_off, items, 20)
    assert pred == [int(v) for v in ei[0] if v >= 0]


def test_end_to_end_sessions_to_topk(oracle):
    """BASELINE config 1 in miniature (test_amazon_filterd.py:485-578): DB = encode(prefix + suffix), queries =
    encode(prefix), cosine top-k — CUDA encoder + CUDA search against oracle encoder + oracle search."""
    import sessionsimilaritysearch_b200 as sss
    from oracle import encoder_oracle as eo
    from sessionsimilaritysearch_b200 import graph, sessions, synth
    in_dim, hidden, n_layers, out_dim, msl = 64, 96, 3, 200, 20
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 21)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    tok = synth.HashTokenizer()
    rng = np.random.default_rng(22)
    full = synth.make_sessions(600, 23)
    pairs = [synth.split_session(s, rng) for s in full]

    def graphs_of(seqs):
        out = []
        for s, t in seqs:
            g = sessions.sequence_to_graph(0, s, t, tok, 20)
            g['query'].x = synth.text_features(g['query'].input_ids, in_dim)
            g['product'].input_ids = synth.text_features(g['product'].input_ids, in_dim)
            out.append(g)
        return out

    db_graphs = graphs_of([(p + s, s) for p, s in pairs])
    q_graphs = graphs_of(pairs[:100])

    def encode(gs, fn):
        return np.concatenate([fn(b) for b in graph.DataLoader(gs, batch_size=200, shuffle=False)], 0)

    cuda_fn = lambda b: enc(b.to("cuda")).cpu().numpy()
    cpu_fn = lambda b: eo.encoder_forward(P, eo.batch_from_pyg(b), n_layers).numpy()
    db_gpu, q_gpu = encode(db_graphs, cuda_fn), encode(q_graphs, cuda_fn)
    db_cpu, q_cpu = encode(db_graphs, cpu_fn), encode(q_graphs, cpu_fn)
    np.testing.assert_allclose(db_gpu, db_cpu, rtol=3e-4, atol=3e-4 * float(np.abs(db_cpu).max()))
    index = sss.build_index(db_gpu, 'cos')
    D, I = index.search(sss.normalize(q_gpu), 20)
    Do, Io = oracle.search_flat(oracle.normalize(db_cpu, 1), oracle.normalize(q_cpu, 1), 20)
    recall = np.mean([len(set(I[r]) & set(Io[r])) / 20.0 for r in range(100)])
    assert recall >= 0.99, recall
    assert np.max(np.abs(D - Do)) < 1e-3
    # on identical embeddings the search itself is exact (d = 200 takes the fp32 scan)
    D2, I2 = oracle.search_flat(oracle.normalize(db_gpu, 1), oracle.normalize(q_gpu, 1), 20)
    assert np.array_equal(I, I2) and np.array_equal(D.view(np.uint32), D2.view(np.uint32))
    # hashed retrieval (fine_tune_ours.py:826-876): sign head -> packed codes -> Hamming top-k
    g = torch.Generator().manual_seed(5)
    W, b = torch.randn(250, out_dim, generator=g) / out_dim ** 0.5, torch.randn(250, generator=g) * 0.1
    head = sss.BinarizeHead(W, b)
    dcodes = sss.pack_sign_bits(head(torch.from_numpy(db_gpu)))
    qcodes = sss.pack_sign_bits(head(torch.from_numpy(q_gpu)))
    bi = sss.IndexBinaryFlat(256)
    bi.add(dcodes)
    Dh, Ih = bi.search(qcodes, 20)
    Dho, Iho = oracle.search_hamming(dcodes.cpu().numpy(), qcodes.cpu().numpy(), 20)
    assert np.array_equal(Dh.cpu().numpy(), Dho) and np.array_equal(Ih.cpu().numpy(), Iho)


def test_config0_ten_thousand_sessions_eval_against_the_oracle_chain():
    """BASELINE configs[0] at its stated size (test_amazon_filterd.py:485-578 on ~10k sessions): database =
    encode(prefix + suffix) of 10,000 synthetic Amazon-filtered-shaped sessions, queries = encode(prefix) of 2,000 of
    them, reference model shape (768 -> 3 x 800 -> 3168 -> 1600), batches of 200, cosine top-100.  The CUDA chain
    (native featuriser -> fused tcgen05 encoder -> tensor-core search) against the CPU oracle chain (encoder oracle ->
    fixed-order search oracle): embeddings inside 3e-4 of the output scale, recall@100 >= 0.999, scores inside 1e-3; and
    on identical embeddings the search is bit-exact."""
    import sessionsimilaritysearch_b200 as sss
    from oracle import encoder_oracle as eo
    from oracle import search_oracle as so
    from sessionsimilaritysearch_b200 import featurize, graph, pipeline, sessions, synth
    in_dim, hidden, n_layers, out_dim, msl = 768, 800, 3, 1600, 20
    n_db, n_q = 10000, 2000
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 31)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    rng = np.random.default_rng(32)
    full = synth.make_sessions(n_db, 33)
    pairs = [synth.split_session(s, rng) for s in full]
    queries = [p for p, _ in pairs[:n_q]]
    vocab = featurize.QueryVocab()
    items = {0}
    for s in full:
        for act in s:
            if act[1] == sessions.SEARCH:
                vocab(act[2])
            else:
                items.add(act[-1])
    item_ids = np.asarray(sorted(items), dtype=np.int64)
    g = torch.Generator().manual_seed(34)
    qf, itf = torch.randn((len(vocab), in_dim), generator=g), torch.randn((len(item_ids), in_dim), generator=g)
    cache = featurize.FeatureCache(qf, item_ids, itf, 0)
    db_gpu, _ = pipeline.encode_sessions(enc, featurize.flatten(full, vocab), cache)
    q_gpu, _ = pipeline.encode_sessions(enc, featurize.flatten(queries, vocab), cache)
    # the oracle chain on the same batches (a batch of 200 is its own graph batch: GATConv's self-loop quirk)
    host_cache = featurize.FeatureCache(qf, item_ids, itf, "cpu")

    def oracle_encode(sess_list):
        flat = featurize.flatten(sess_list, vocab)
        out = []
        for lo in range(0, len(flat), 200):
            b = featurize.featurize_batch(flat.slice(lo, min(len(flat), lo + 200)), host_cache)
            out.append(eo.encoder_forward(P, eo.batch_from_pyg(b), n_layers))
        return torch.cat(out).numpy()

    db_cpu = oracle_encode(full[:2000])           # the encoder oracle on a 2,000-session slice (time)
    scale = float(np.abs(db_cpu).max())
    np.testing.assert_allclose(db_gpu[:2000].cpu().numpy(), db_cpu, rtol=3e-4, atol=3e-4 * scale)
    index = sss.build_index(db_gpu, 'cos')
    D, I = index.search(sss.normalize(q_gpu), 100)
    D, I = D.cpu().numpy(), I.cpu().numpy()
    # search parity on the very same embeddings: ids and scores bit for bit
    Do, Io = so.search_flat(so.normalize(db_gpu.cpu().numpy(), 1), so.normalize(q_gpu.cpu().numpy(), 1)[:64], 100)
    assert np.array_equal(I[:64], Io) and np.array_equal(D[:64].view(np.uint32), Do.view(np.uint32))
    # the query prefix must find its own session
    assert float((I[:, 0] == np.arange(n_q)).mean()) >= 0.4   # (random-init weights, uniformly random split point)
    # whole-chain agreement where both chains were run: oracle embeddings of the first 2,000 database sessions
    q_cpu = oracle_encode(queries[:200])
    sub = sss.build_index(db_gpu[:2000].contiguous(), 'cos')
    Ds, Is = sub.search(sss.normalize(q_gpu[:200].contiguous()), 100)
    Dc, Ic = so.search_flat(so.normalize(db_cpu, 1), so.normalize(q_cpu, 1), 100)
    Is, Ds = Is.cpu().numpy(), Ds.cpu().numpy()
    recall = np.mean([len(set(Is[r]) & set(Ic[r])) / 100.0 for r in range(200)])
    assert recall >= 0.999, recall
    assert np.max(np.abs(Ds - Dc)) < 1e-3
