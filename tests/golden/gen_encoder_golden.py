"""Encoder + featuriser goldens from the reference's OWN code, run in the build container.

    python tests/golden/gen_encoder_golden.py     (needs /root/reference; not needed at test time)

With oracle/pyg_shim on sys.path (a plain-torch stand-in for the absent torch_geometric 2.0.4), the reference's
model/gnn.py, model/model.py and util_amazon_filtered.py import and run unmodified:
  * sequence_to_graph (util_amazon_filtered.py:98-230) on seeded synthetic sessions -> graph tensors;
  * UnifyPoolingGraphLevelEncoder (model/model.py:263-351) built from HeteroGGNN (model/gnn.py:43-81) and
    PositionalAttentionPooling (model/gnn.py:183-217), seeded weights loaded into its state_dict, the frozen text
    embedder replaced by an identity over precomputed features (its weights are private) -> [B, out];
  * BinarizeHead (model/model.py:105-138) in eval mode -> codes.
Outputs: tests/golden/encoder_golden_<cfg>.npz, tests/golden/graphs_golden.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)
sys.modules.setdefault("Levenshtein", types.ModuleType("Levenshtein"))

import encoder_common as ec  # noqa: E402
import util_amazon_filtered as ref_util  # noqa: E402  (the reference's)
from model.gnn import HeteroGGNN, PositionalAttentionPooling  # noqa: E402
from model.model import BinarizeHead, UnifyPoolingGraphLevelEncoder  # noqa: E402
from model.NodeEmbedding import NodeAsinEmbedding  # noqa: E402
from torch_geometric.loader import DataLoader  # noqa: E402  (the shim)


class FeatureEmbedder(torch.nn.Module):
    """stands in for PretrainedQAEAEncoder(None) (model/NodeEmbedding.py:100-125): features arrive precomputed"""

    def forward(self, input_ids, token_type_ids, attention_mask, get_token=False):
        return (input_ids, None) if get_token else input_ids


def graph_tensors(graphs):
    out = {}
    for i, g in enumerate(graphs):
        for nt in ("query", "product"):
            for a in ("pos_emb_id",) + (("cnt", "x", "last_click_mask") if nt == "product" else ("mask",)):
                out["g%d_%s_%s" % (i, nt, a)] = np.asarray(g[nt][a])
        out["g%d_query_tokens" % i] = np.asarray(g["query"].input_ids)
        for name, et in (("qp", ec.EDGE_QP), ("pq", ec.EDGE_PQ), ("pp", ec.EDGE_PP)):
            out["g%d_%s" % (i, name)] = np.asarray(g[et].edge_index)
        out["g%d_pp_w" % i] = np.asarray(g[ec.EDGE_PP].edge_weight)
        out["g%d_target_y" % i] = np.asarray(g["product_target"].y)
        out["g%d_text_tokens" % i] = np.asarray(g["text"].input_ids)
    return out


def main():
    torch.manual_seed(0)
    # ---- featuriser golden: reference graphs BEFORE the float features overwrite x / input_ids
    from sessionsimilaritysearch_b200 import synth
    tok = synth.HashTokenizer()
    sess = synth.make_sessions(24, 5)
    sess[3] = [a for a in sess[3] if a[1] == 's'] or sess[3]
    ref_graphs = [ref_util.sequence_to_graph(0, s, s[len(s) // 2:], tok, 20) for s in sess]
    ig = [ref_util.sequence_to_graph(0, s, s[len(s) // 2:], tok, 20, True) for s in sess[:6]]
    gt = graph_tensors(ref_graphs)
    gt.update({"ig_" + k: v for k, v in graph_tensors(ig).items()})
    for i, g in enumerate(ref_graphs):
        gt["g%d_product_tokens" % i] = np.asarray(g["product"].input_ids)
    np.savez_compressed(os.path.join(HERE, "graphs_golden.npz"), **gt)
    print("wrote graphs_golden.npz with", len(gt), "arrays")

    # ---- encoder goldens
    for name, in_dim, hidden, n_layers, out_dim, msl, n_sess, seed in ec.CONFIGS:
        _, graphs = ec.make_graphs(n_sess, in_dim, seed, ref_util.sequence_to_graph)
        gnn = HeteroGGNN(hidden, n_layers, graphs[0])
        node_dim = in_dim + n_layers * hidden
        pooling = PositionalAttentionPooling(node_dim, node_dim, out_dim, msl)
        enc = UnifyPoolingGraphLevelEncoder(FeatureEmbedder(), NodeAsinEmbedding(16, 8), gnn, pooling, None,
                                            use_id_embedding=False)
        enc.eval()
        batch = next(iter(DataLoader(graphs, batch_size=len(graphs), shuffle=False)))
        batch["product"].x = batch["product"].x % 16  # ids only feed the discarded ASIN embedding
        with torch.no_grad():
            enc(batch)  # materialises the lazy (-1) linears
            P = ec.make_params(in_dim, hidden, n_layers, out_dim, msl, seed)
            sd = enc.state_dict()
            for k, v in P.items():
                assert k in sd and tuple(sd[k].shape) == tuple(v.shape), (k, v.shape, sd.get(k, torch.empty(0)).shape)
                sd[k].copy_(v)
            out, nodes = enc(batch, get_node=True)
            head = BinarizeHead(out_dim, 250 if name == "full" else 20, None)
            head.eval()
            g = torch.Generator().manual_seed(seed + 1)
            head.lin1.weight.copy_(torch.randn(head.lin1.weight.shape, generator=g) / out_dim ** 0.5)
            head.lin1.bias.copy_(torch.randn(head.lin1.bias.shape, generator=g) * 0.1)
            codes = head(out)
        np.savez_compressed(os.path.join(HERE, "encoder_golden_%s.npz" % name), out=out.numpy(),
                            node_query=nodes["query"].numpy()[:, -hidden:], node_product=nodes["product"].numpy()[:, -hidden:],
                            codes=codes.numpy(), head_w=head.lin1.weight.detach().numpy(), head_b=head.lin1.bias.detach().numpy())
        print(name, "out", tuple(out.shape), "abs mean %.4f" % out.abs().mean().item(), "codes", tuple(codes.shape))


if __name__ == "__main__":
    main()
