"""CPU tests of the encoder-side oracle and host logic against goldens produced by the reference's own code
(tests/golden/gen_encoder_golden.py): featuriser layout, batching, functional encoder restatement, hash head."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import encoder_common as ec  # noqa: E402

from oracle import encoder_oracle as eo  # noqa: E402
from sessionsimilaritysearch_b200 import graph, sessions, synth  # noqa: E402

GG = np.load(os.path.join(HERE, "golden", "graphs_golden.npz"))


def _check_graph(prefix, g):
    for nt in ("query", "product"):
        for a in ("pos_emb_id",) + (("cnt", "x", "last_click_mask") if nt == "product" else ("mask",)):
            assert np.array_equal(GG["%s_%s_%s" % (prefix, nt, a)], np.asarray(g[nt][a])), (prefix, nt, a)
    assert np.array_equal(GG[prefix + "_query_tokens"], np.asarray(g["query"].input_ids))
    for name, et in (("qp", ec.EDGE_QP), ("pq", ec.EDGE_PQ), ("pp", ec.EDGE_PP)):
        assert np.array_equal(GG["%s_%s" % (prefix, name)], np.asarray(g[et].edge_index)), (prefix, name)
    assert np.array_equal(GG[prefix + "_pp_w"], np.asarray(g[ec.EDGE_PP].edge_weight))
    assert np.array_equal(GG[prefix + "_target_y"], np.asarray(g["product_target"].y))
    assert np.array_equal(GG[prefix + "_text_tokens"], np.asarray(g["text"].input_ids))


def test_sequence_to_graph_matches_the_reference():
    tok = synth.HashTokenizer()
    sess = synth.make_sessions(24, 5)
    sess[3] = [a for a in sess[3] if a[1] == 's'] or sess[3]
    for i, s in enumerate(sess):
        g = sessions.sequence_to_graph(0, s, s[len(s) // 2:], tok, 20)
        _check_graph("g%d" % i, g)
        assert np.array_equal(GG["g%d_product_tokens" % i], np.asarray(g["product"].input_ids))
        assert g["ori_seq"] == (s, s[len(s) // 2:])
    for i, s in enumerate(sess[:6]):  # ignore_query=True arm (util_amazon_filtered.py:101-103)
        _check_graph("ig_g%d" % i, sessions.sequence_to_graph(0, s, s[len(s) // 2:], tok, 20, True))


def test_itemless_session_gets_the_placeholder_product():
    tok = synth.HashTokenizer()
    s = [(0, 's', 'a query', None, None, None, None, 0), (1, 's', None, None, None, None, None, 0)]
    g = sessions.sequence_to_graph(7, s, s[:1], tok, 20)
    assert g["product"].x.tolist() == [0] and g["product"].cnt.tolist() == [1] and g["product"].pos_emb_id.tolist() == [0]
    assert g["query"].x.shape[0] == 3 and g["query"].pos_emb_id.tolist() == [2, 1, 0]
    assert g[ec.EDGE_QP].edge_index.shape == (2, 0) and g[ec.EDGE_PP].edge_index.shape == (2, 0)
    assert g["idx"].idx == 7


def test_collate_offsets_and_batch_vectors():
    tok = synth.HashTokenizer()
    gs = [sessions.sequence_to_graph(0, s, s[:1], tok, 20) for s in synth.make_sessions(5, 3)]
    b = graph.collate(gs)
    nq = [g["query"].x.shape[0] for g in gs]
    npd = [g["product"].x.shape[0] for g in gs]
    assert b["query"].batch.tolist() == sum([[i] * n for i, n in enumerate(nq)], [])
    assert b["product"].batch.tolist() == sum([[i] * n for i, n in enumerate(npd)], [])
    assert list(b.edge_index_dict.keys()) == [ec.EDGE_QP, ec.EDGE_PQ, ec.EDGE_PP]
    oq = np.concatenate([[0], np.cumsum(nq)])
    op = np.concatenate([[0], np.cumsum(npd)])
    exp = torch.cat([g[ec.EDGE_QP].edge_index + torch.tensor([[oq[i]], [op[i]]]) for i, g in enumerate(gs)], 1)
    assert torch.equal(b[ec.EDGE_QP].edge_index, exp)
    exp = torch.cat([g[ec.EDGE_PP].edge_index + int(op[i]) for i, g in enumerate(gs)], 1)
    assert torch.equal(b[ec.EDGE_PP].edge_index, exp)
    assert b["product"].cnt.sum() == b["product"].pos_emb_id.numel() and b.num_graphs == 5
    loader = graph.DataLoader(gs, batch_size=2, shuffle=False)
    assert [x.num_graphs for x in loader] == [2, 2, 1]


@pytest.mark.parametrize("cfg", ec.CONFIGS, ids=[c[0] for c in ec.CONFIGS])
def test_encoder_oracle_matches_the_reference_model(cfg):
    name, in_dim, hidden, n_layers, out_dim, msl, n_sess, seed = cfg
    gold = np.load(os.path.join(HERE, "golden", "encoder_golden_%s.npz" % name))
    _, graphs = ec.make_graphs(n_sess, in_dim, seed, sessions.sequence_to_graph)
    batch = eo.batch_from_pyg(graph.collate(graphs))          # this repo's featuriser + batcher
    P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, seed)
    out, zq, zp = eo.encoder_forward(P, batch, n_layers, return_nodes=True)
    tol = dict(rtol=2e-4, atol=2e-4 * float(np.abs(gold["out"]).max()))
    np.testing.assert_allclose(zq[:, -hidden:].numpy(), gold["node_query"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(zp[:, -hidden:].numpy(), gold["node_product"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(out.numpy(), gold["out"], **tol)
    codes = eo.binarize_head_eval(torch.from_numpy(gold["out"]), torch.from_numpy(gold["head_w"]),
                                  torch.from_numpy(gold["head_b"]))
    assert np.array_equal(codes.numpy(), gold["codes"]) and set(np.unique(gold["codes"])) <= {-1.0, 0.0, 1.0}
    # float64 run of the same restatement bounds the fp32 noise of both
    P64 = {k: v.double() for k, v in P.items()}
    out64 = eo.encoder_forward(P64, batch, n_layers)
    assert float((out64 - torch.from_numpy(gold["out"]).double()).abs().max()) < tol["atol"]
