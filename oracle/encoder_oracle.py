"""CPU restatement (plain torch, fp32 or fp64) of the reference's session encoder forward.

TEST INFRASTRUCTURE ONLY.  Follows, line by line:
    UnifyPoolingGraphLevelEncoder.forward   model/model.py:279-351   (use_id_embedding=False arm, no node masks)
    HeteroGGNN.forward                      model/gnn.py:64-81       (add_input_feat=True)
    PositionalAttentionPooling.forward      model/gnn.py:193-217
    BinarizeHead.forward (eval, mlp=None)   model/model.py:117-138
and the PyG 2.0.4 layers they call (GATConv / GatedGraphConv / HeteroConv / global_mean_pool), whose
semantics are recalled — see oracle/pyg_shim/torch_geometric/nn/__init__.py and SURVEY.md Appendix A.
Pinned by tests/golden/encoder_golden_*.npz, which were produced by running the reference's OWN
model/gnn.py + model/model.py over the shim (tests/golden/gen_encoder_golden.py).

Parameters are a dict keyed like the reference state_dict:
    gnn.convs.{l}.convs.query__clicks__product.{lin_src.weight,lin_dst.weight,att_src,att_dst,bias}
    gnn.convs.{l}.convs.product__clicked by__query.{...}
    gnn.convs.{l}.convs.product__to__product.{weight,rnn.weight_ih,rnn.weight_hh,rnn.bias_ih,rnn.bias_hh}
    pooling.{query_lin,product_lin,node_emb_lin}.{weight,bias}, pooling.positional_emb.weight,
    pooling.coarse_rep_lin.weight, pooling.att_lin.weight
The batch is a dict with the attributes the reference reads (model/model.py:283-286,317; model/gnn.py:199-206).
"""
import torch
import torch.nn.functional as F

EDGE_QP = "query__clicks__product"
EDGE_PQ = "product__clicked by__query"
EDGE_PP = "product__to__product"


def _segment_sum(src, index, n):
    out = src.new_zeros((n,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def gat(P, prefix, x_src, x_dst, src, dst):
    """GATConv((-1,-1), C), heads=1, add_self_loops=True on a bipartite edge set (batch-global indices)"""
    xs = x_src @ P[prefix + "lin_src.weight"].T
    xd = x_dst @ P[prefix + "lin_dst.weight"].T
    a_s = (xs * P[prefix + "att_src"].view(1, -1)).sum(-1)
    a_d = (xd * P[prefix + "att_dst"].view(1, -1)).sum(-1)
    keep = src != dst
    n_loop = min(x_src.shape[0], x_dst.shape[0])
    loop = torch.arange(n_loop, dtype=src.dtype)
    j = torch.cat([src[keep], loop])
    i = torch.cat([dst[keep], loop])
    e = F.leaky_relu(a_s[j] + a_d[i], 0.2)
    n_dst = x_dst.shape[0]
    emax = torch.full((n_dst,), float("-inf"), dtype=e.dtype).scatter_reduce(0, i, e, reduce="amax", include_self=True)
    ex = (e - emax[i]).exp()
    alpha = ex / (_segment_sum(ex, i, n_dst)[i] + 1e-16)
    out = _segment_sum(xs[j] * alpha[:, None], i, n_dst)
    return out + P[prefix + "bias"]


def ggc(P, prefix, x, src, dst):
    """GatedGraphConv(C, 1) without edge weights + torch.nn.GRUCell"""
    C = P[prefix + "weight"].shape[-1]
    if x.shape[1] > C:
        raise ValueError("input wider than GatedGraphConv channels")
    if x.shape[1] < C:
        x = torch.cat([x, x.new_zeros(x.shape[0], C - x.shape[1])], 1)
    m = x @ P[prefix + "weight"][0]
    a = _segment_sum(m[src], dst, x.shape[0])
    gi = a @ P[prefix + "rnn.weight_ih"].T + P[prefix + "rnn.bias_ih"]
    gh = x @ P[prefix + "rnn.weight_hh"].T + P[prefix + "rnn.bias_hh"]
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * x


def encoder_forward(P, batch, n_layers=3, return_nodes=False):
    """-> graph embeddings [B, out_dim]"""
    dt = P["pooling.att_lin.weight"].dtype
    xq = batch["x_query"].to(dt)
    xp = batch["x_product"].to(dt)
    q_list, p_list = [xq], [xp]
    for l in range(n_layers):
        pre = "gnn.convs.%d.convs." % l
        hq, hp = q_list[-1], p_list[-1]
        # HeteroConv iterates edge_index_dict in insertion order: q->p, p->q, p->p (util_amazon_filtered.py:194-195,217)
        g_p = gat(P, pre + EDGE_QP + ".", hq, hp, batch["qp_src"], batch["qp_dst"])
        g_q = gat(P, pre + EDGE_PQ + ".", hp, hq, batch["pq_src"], batch["pq_dst"])
        r_p = ggc(P, pre + EDGE_PP + ".", hp, batch["pp_src"], batch["pp_dst"])
        p_list.append(torch.relu(torch.stack([g_p, r_p], 0).sum(0)))
        q_list.append(torch.relu(g_q))
    zq = torch.cat(q_list, 1)
    zp = torch.cat(p_list, 1)
    # PositionalAttentionPooling (model/gnn.py:193-217)
    pe = P["pooling.positional_emb.weight"]
    uq = zq @ P["pooling.query_lin.weight"].T + P["pooling.query_lin.bias"]
    up = zp @ P["pooling.product_lin.weight"].T + P["pooling.product_lin.bias"]
    uq = torch.tanh(torch.cat([uq, pe[batch["query_pos"]]], 1))
    up = torch.repeat_interleave(up, batch["product_cnt"], dim=0)
    up = torch.tanh(torch.cat([up, pe[batch["product_pos"]]], 1))
    pb = torch.repeat_interleave(batch["product_batch"], batch["product_cnt"], dim=0)
    nodes = torch.cat([up, uq], 0)
    nb = torch.cat([pb, batch["query_batch"]], 0)
    B = int(nb.max().item()) + 1
    cnt = _segment_sum(torch.ones_like(nb, dtype=dt), nb, B).clamp(min=1)[:, None]
    coarse = (_segment_sum(nodes, nb, B) / cnt)[nb]
    a = nodes @ P["pooling.node_emb_lin.weight"].T + P["pooling.node_emb_lin.bias"]
    b = coarse @ P["pooling.coarse_rep_lin.weight"].T
    att = torch.sigmoid(a + b) @ P["pooling.att_lin.weight"].T
    out = _segment_sum(nodes * att, nb, B) / cnt
    if return_nodes:
        return out, zq, zp
    return out


def binarize_head_eval(x, W, b):
    """BinarizeHead.forward in eval mode with mlp=None: numerically sign(x W^T + b) (model/model.py:137)"""
    out = x @ W.T + b
    return (torch.sign(out) - torch.tanh(out)).detach() + torch.tanh(out)


def batch_from_pyg(data):
    """attributes of a (shim or real) PyG HeteroDataBatch -> the dict encoder_forward reads; text features
    are expected in data['query'].x and data['product'].input_ids (what the embedder is fed, model/model.py:283,286)"""
    e = data.edge_index_dict
    return {
        "x_query": data["query"].x, "x_product": data["product"].input_ids,
        "query_batch": data["query"].batch, "product_batch": data["product"].batch,
        "query_pos": data["query"].pos_emb_id, "product_cnt": data["product"].cnt,
        "product_pos": data["product"].pos_emb_id,
        "qp_src": e[("query", "clicks", "product")][0], "qp_dst": e[("query", "clicks", "product")][1],
        "pq_src": e[("product", "clicked by", "query")][0], "pq_dst": e[("product", "clicked by", "query")][1],
        "pp_src": e[("product", "to", "product")][0], "pp_dst": e[("product", "to", "product")][1],
    }
