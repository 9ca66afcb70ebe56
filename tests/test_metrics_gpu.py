"""GPU parity tests of the rows around the search path: evaluation metrics on the retrieved ids,
the neighbour item vote against the fixture the reference's own function produced, the in-batch cosine matrix (a15),
the text embedder's masked mean (a3) and the standalone gnn() / pooling() / get_node entry points (8b)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import encoder_common as ec  # noqa: E402
from test_metrics_cpu import golden_data  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("sim", ["all_jaccard", "cur_jaccard", "all_product_type_score"])
def test_pair_scores_match_reference_generated_scores(sim):
    import sessionsimilaritysearch_b200 as sss
    z, test, train = golden_data()
    gt = sss.score_matrix(z["I"], test, train, sim)
    assert np.array_equal(gt.view(np.uint32), z["gt_" + sim].view(np.uint32))
    assert np.float32(sss.get_ave_score(z["I"], test, train, sim)) == z["mean_" + sim]
    # single-pair form of the call
    from oracle import metrics_oracle as mo
    for i, j in ((0, 0), (3, 7), (11, 2)):
        r = train[int(z["I"][i, j])]
        assert np.float32(sss.get_score(test[i], (r, []), sim)) == np.float32(mo.get_score(test[i], (r, []), sim))


def test_pair_scores_larger_batch_against_the_oracle():
    import sessionsimilaritysearch_b200 as sss
    from oracle import metrics_oracle as mo
    from sessionsimilaritysearch_b200 import synth
    rng = np.random.default_rng(5)
    test = [synth.split_session(s, rng) for s in synth.make_sessions(300, 6)]
    train = synth.make_sessions(5000, 7)
    I = rng.integers(0, len(train), size=(300, 100)).astype(np.int64)
    for sim in ("all_jaccard", "cur_jaccard", "all_product_type_score"):
        got = sss.score_matrix(torch.from_numpy(I).cuda(), test, train, sim)
        exp = mo.score_matrix(I, test, train, sim)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), sim


@pytest.mark.parametrize("name", ["small", "heavy_ties", "equal_weights", "ref_call_500", "ref_call_500_ties"])
def test_item_vote_matches_the_reference_function(oracle, name):
    """fixture = outputs of the reference's get_prediction_by_knn (float64 sums in arrival order, stable sort)"""
    import sessionsimilaritysearch_b200 as sss
    z = np.load(os.path.join(GOLD, "vote_golden.npz"))
    item_off, items, D, I, K = (z[name + "_item_off"], z[name + "_items"], z[name + "_D"], z[name + "_I"],
                                int(z[name + "_K"]))
    lists = sss.ItemLists([items[item_off[i]:item_off[i + 1]] for i in range(len(item_off) - 1)])
    oi, ow = sss.item_vote(D, I, lists, K)
    assert np.array_equal(oi.cpu().numpy()[0], z[name + "_expected"])
    ei, ew = oracle.item_vote(D, I, item_off, items, K)
    assert np.array_equal(oi.cpu().numpy(), ei) and np.array_equal(ow.cpu().numpy().view(np.uint32), ew.view(np.uint32))
    # many queries at once, the reference's call shape (sample_size 500, K 20)
    if name == "ref_call_500":
        rng = np.random.default_rng(3)
        Ib = np.stack([rng.choice(len(item_off) - 1, size=500, replace=False) for _ in range(64)]).astype(np.int64)
        Db = np.sort(rng.uniform(0.2, 0.99, size=(64, 500)).astype(np.float32), axis=1)[:, ::-1].copy()
        Db[5] = 0.5
        oi, ow = sss.item_vote(Db, Ib, lists, K)
        ei, ew = oracle.item_vote(Db, Ib, item_off, items, K)
        assert np.array_equal(oi.cpu().numpy(), ei)
        assert np.array_equal(ow.cpu().numpy().view(np.uint32), ew.view(np.uint32))


def test_item_vote_duplicate_items_and_limits(oracle):
    import sessionsimilaritysearch_b200 as sss
    # the same item twice inside one neighbour's list, long lists (> 32 items), padded neighbours
    rng = np.random.default_rng(9)
    lens = rng.integers(1, 80, size=200)
    item_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    items = rng.integers(0, 150, size=item_off[-1]).astype(np.int64)
    lists = sss.ItemLists([items[item_off[i]:item_off[i + 1]] for i in range(200)])
    I = np.stack([rng.permutation(200)[:40] for _ in range(7)]).astype(np.int64)
    I[2, 30:] = -1
    D = rng.choice(np.array([0.25, 0.5, 1.0], np.float32), size=(7, 40))
    oi, ow = sss.item_vote(D, I, lists, 30)
    ei, ew = oracle.item_vote(D, I, item_off, items, 30)
    assert np.array_equal(oi.cpu().numpy(), ei) and np.array_equal(ow.cpu().numpy().view(np.uint32), ew.view(np.uint32))
    with pytest.raises(RuntimeError):
        sss.item_vote(D, I, lists, 300)      # K > 256


def test_cosine_matrix_matches_torch():
    import sessionsimilaritysearch_b200 as sss
    g = torch.Generator().manual_seed(1)
    for na, nb, d in ((50, 50, 1600), (7, 33, 250), (1, 1, 3)):
        a, b = torch.randn(na, d, generator=g), torch.randn(nb, d, generator=g)
        a[0] = 0                                    # F.normalize's 1e-12 floor
        ref = torch.nn.functional.normalize(a.double()) @ torch.nn.functional.normalize(b.double()).T
        got = sss.cosine_matrix(a.cuda(), b.cuda()).cpu()
        assert got.shape == (na, nb)
        assert float((got.double() - ref).abs().max()) <= 2e-6     # fp32 tolerance of the reference's own matmul
        ref32 = torch.nn.functional.normalize(a) @ torch.nn.functional.normalize(b).T
        assert float((got - ref32).abs().max()) <= 2e-6


def test_masked_mean_pool_matches_the_reference_expression():
    import sessionsimilaritysearch_b200 as sss
    g = torch.Generator().manual_seed(2)
    tok = torch.randn(37, 20, 768, generator=g)
    mask = (torch.rand(37, 20, generator=g) < 0.6).long()
    mask[:, 0] = 1
    ref = torch.sum(tok * mask.unsqueeze(-1), dim=1) / torch.sum(mask, dim=1).view(-1, 1)   # model/NodeEmbedding.py:113
    out, tk = sss.masked_mean_pool(tok.cuda(), mask.cuda(), get_token=True)
    assert tk.shape == tok.shape and out.shape == (37, 768)
    assert float((out.cpu() - ref).abs().max()) <= 1e-6 * float(ref.abs().max()) + 1e-7


def test_standalone_gnn_pooling_and_get_node_agree_with_the_forward():
    """gnn(x_dict, edge_index_dict), pooling(node_emb, data) and encoder(data, get_node=True): the stages run on their
    own must reproduce the one-call forward bit for bit, and the node embeddings must equal the oracle's."""
    import sessionsimilaritysearch_b200 as sss
    from oracle import encoder_oracle as eo
    from sessionsimilaritysearch_b200 import graph, sessions
    in_dim, hidden, n_layers, out_dim, msl = 24, 32, 3, 60, 20
    _, graphs = ec.make_graphs(12, in_dim, 7, sessions.sequence_to_graph)
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, 7)
    enc = sss.SessionEncoder(P, in_dim=in_dim, hidden=hidden, n_layers=n_layers, out_dim=out_dim, max_seq_len=msl)
    data = graph.collate(graphs).to("cuda")
    out = enc(data)
    out2, nodes = enc(data, get_node=True)
    assert torch.equal(out, out2)
    assert nodes["query"].shape == (data["query"].x.shape[0], in_dim + n_layers * hidden)
    out3, tok = enc(data, get_token=True)
    assert tok == {} and torch.equal(out3, out)
    out4, nodes4, tok4 = enc(data, get_node=True, get_token=True)
    assert tok4 == {} and torch.equal(nodes4["product"], nodes["product"])
    z = enc.gnn({"query": data["query"].x, "product": data["product"].input_ids}, data.edge_index_dict)
    assert torch.equal(z["query"], nodes["query"]) and torch.equal(z["product"], nodes["product"])
    z_no = enc.gnn({"query": data["query"].x, "product": data["product"].input_ids}, data.edge_index_dict,
                   add_input_feat=False)
    assert torch.equal(z_no["query"], nodes["query"][:, in_dim:])
    assert torch.equal(enc.pooling(z, data), out)
    ref_out, ref_zq, ref_zp = eo.encoder_forward(P, eo.batch_from_pyg(graph.collate(graphs)), n_layers, return_nodes=True)
    for got, r in ((nodes["query"], ref_zq), (nodes["product"], ref_zp), (out, ref_out)):
        assert torch.allclose(got.cpu(), r, rtol=2e-4, atol=2e-4 * float(r.abs().max()))
