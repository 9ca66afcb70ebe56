"""Shared by gen_encoder_golden.py (build container, needs /root/reference) and the tests (anywhere):
deterministic synthetic sessions, weights and text features, so that goldens can be committed as small files
(inputs and weights are regenerated from seeds; only the reference's outputs are stored)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sessionsimilaritysearch_b200 import synth  # noqa: E402

EDGE_QP = ("query", "clicks", "product")
EDGE_PQ = ("product", "clicked by", "query")
EDGE_PP = ("product", "to", "product")

# (name, in_dim, hidden, n_layers, out_dim, max_seq_len, n_sessions, seed)
CONFIGS = [
    ("small", 24, 32, 3, 60, 20, 12, 7),
    ("full", 768, 800, 3, 1600, 20, 6, 11),       # the reference's shapes (pretrain_filtered_amazon.py:262-287)
]


def param_shapes(in_dim, hidden, n_layers, out_dim, max_seq_len):
    """state_dict keys and shapes of UnifyPoolingGraphLevelEncoder's live parameters"""
    shapes = {}
    for l in range(n_layers):
        cin = in_dim if l == 0 else hidden
        for et in ("query__clicks__product", "product__clicked by__query"):
            pre = "gnn.convs.%d.convs.%s." % (l, et)
            shapes[pre + "lin_src.weight"] = (hidden, cin)
            shapes[pre + "lin_dst.weight"] = (hidden, cin)
            shapes[pre + "att_src"] = (1, 1, hidden)
            shapes[pre + "att_dst"] = (1, 1, hidden)
            shapes[pre + "bias"] = (hidden,)
        pre = "gnn.convs.%d.convs.product__to__product." % l
        shapes[pre + "weight"] = (1, hidden, hidden)
        shapes[pre + "rnn.weight_ih"] = (3 * hidden, hidden)
        shapes[pre + "rnn.weight_hh"] = (3 * hidden, hidden)
        shapes[pre + "rnn.bias_ih"] = (3 * hidden,)
        shapes[pre + "rnn.bias_hh"] = (3 * hidden,)
    node_dim = in_dim + n_layers * hidden
    shapes["pooling.query_lin.weight"] = (out_dim - max_seq_len, node_dim)
    shapes["pooling.query_lin.bias"] = (out_dim - max_seq_len,)
    shapes["pooling.product_lin.weight"] = (out_dim - max_seq_len, node_dim)
    shapes["pooling.product_lin.bias"] = (out_dim - max_seq_len,)
    shapes["pooling.positional_emb.weight"] = (max_seq_len, max_seq_len)
    shapes["pooling.node_emb_lin.weight"] = (out_dim, out_dim)
    shapes["pooling.node_emb_lin.bias"] = (out_dim,)
    shapes["pooling.coarse_rep_lin.weight"] = (out_dim, out_dim)
    shapes["pooling.att_lin.weight"] = (1, out_dim)
    return shapes


def make_params(in_dim, hidden, n_layers, out_dim, max_seq_len, seed):
    """seeded weights; scale ~ 1/sqrt(fan_in) keeps activations O(1) through three layers"""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for k, shp in param_shapes(in_dim, hidden, n_layers, out_dim, max_seq_len).items():
        fan_in = shp[-1] if len(shp) > 1 else hidden
        scale = 1.0 / (fan_in ** 0.5)
        if k.endswith("positional_emb.weight") or "att_" in k:
            scale = 0.5
        if k.endswith("bias") or "bias_" in k:
            scale = 0.1
        P[k] = torch.randn(shp, generator=g) * scale
    return P


def make_graphs(n_sessions, in_dim, seed, graph_fn, tokenizer=None):
    """sessions -> graphs via graph_fn (the reference's or this repo's sequence_to_graph), with the text
    features placed where the encoder reads them (data['query'].x, data['product'].input_ids)"""
    tok = tokenizer or synth.HashTokenizer()
    sessions = synth.make_sessions(n_sessions, seed)
    sessions[1] = [a for a in sessions[1] if a[1] == 's'] or sessions[1]      # an item-less session
    graphs = []
    for s in sessions:
        g = graph_fn(0, s, s[:1], tok, 20)
        g['query'].x = synth.text_features(g['query'].input_ids, in_dim)
        g['product'].input_ids = synth.text_features(g['product'].input_ids, in_dim)
        graphs.append(g)
    return sessions, graphs
