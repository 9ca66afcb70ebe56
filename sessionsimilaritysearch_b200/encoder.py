"""Session encoder façade: the call surface of the reference's UnifyPoolingGraphLevelEncoder
(model/model.py:263-351) on top of sss_encoder_* (csrc/encoder.cu).

    enc = SessionEncoder(state_dict_like, in_dim=768, hidden=800, n_layers=3, out_dim=1600, max_seq_len=20)
    emb = enc(data)                      # data: a SessionBatch / PyG HeteroDataBatch on the GPU -> [B, out_dim]

`data['query'].x` and `data['product'].input_ids` carry the text features [N, in_dim] — the output of the frozen text
embedder (model/NodeEmbedding.py:112-125), which is where the CUDA scope starts.
"""
import ctypes

import torch

from . import _lib
from ._lib import check

EDGE_QP = ("query", "clicks", "product")
EDGE_PQ = ("product", "clicked by", "query")
EDGE_PP = ("product", "to", "product")


def _i64(t, dev):
    return t.to(device=dev, dtype=torch.int64).contiguous()


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class SessionEncoder:
    def __init__(self, params, in_dim=768, hidden=800, n_layers=3, out_dim=1600, max_seq_len=20, device=None,
                 math="bf16x3"):
        self._lib = _lib.load()
        self.device = _lib.current_device() if device is None else int(device)
        self.in_dim, self.hidden, self.n_layers, self.out_dim, self.max_seq_len = in_dim, hidden, n_layers, out_dim, max_seq_len
        shape = _lib.EncoderShape(in_dim, hidden, n_layers, out_dim, max_seq_len)
        h = ctypes.c_void_p()
        check(self._lib.sss_encoder_create(ctypes.byref(h), self.device, ctypes.byref(shape)))
        self._h = h
        self._pending = []
        self._keep = None
        self.load_state_dict(params)
        # dense linears: this library's split-bf16 tcgen05 GEMM (2.7e-5 of the output scale from float64), the only
        # arithmetic of the C ABI
        self.set_math(math)

    @classmethod
    def from_module(cls, module, **kw):
        """build from a reference-shaped torch module (or anything with .state_dict())"""
        return cls(module.state_dict(), **kw)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.sss_encoder_destroy(h)

    def load_state_dict(self, params):
        """accepts the reference's key names; unrelated keys (text model, dead heads) are ignored"""
        n = 0
        for k, v in params.items():
            if not (k.startswith("gnn.convs.") or k.startswith("pooling.")):
                continue
            t = v.detach().to(torch.float32).contiguous().cpu()
            check(self._lib.sss_encoder_set_param(self._h, k.encode(), t.data_ptr(), t.numel(), 0, None))
            n += 1
        return n

    def set_math(self, math):
        """'bf16x3': this library's tcgen05 GEMM on split-bf16 operands (fp32-level accuracy on the tensor cores);
        the cuBLAS arithmetics of earlier versions ('fp32', 'bf16x9') raise RuntimeError"""
        check(self._lib.sss_encoder_set_math(self._h, {"fp32": 0, "bf16x9": 1, "bf16x3": 2}[math]))
        return self

    @property
    def math(self):
        return ("fp32", "bf16x9", "bf16x3")[int(self._lib.sss_encoder_get_math(self._h))]

    @property
    def launches(self):
        """kernels launched by the last forward"""
        return int(self._lib.sss_encoder_stat(self._h, 0))

    @property
    def host_ns(self):
        """host nanoseconds the last forward spent enqueueing its kernels (inside the C call)"""
        return int(self._lib.sss_encoder_stat(self._h, 1))

    def eval(self):
        return self

    def to(self, device):
        return self

    def _batch(self, data, xq, xp):
        """the sss_graph_batch_t of a SessionBatch / PyG batch (+ the tensors that must outlive the call)"""
        dev = torch.device("cuda", self.device)
        ei = data.edge_index_dict
        # rows first: a [2, E] column slice of a larger edge list is not contiguous, its two rows are
        rows = [_i64(ei[key][r], dev) for key in (EDGE_QP, EDGE_PQ, EDGE_PP) for r in (0, 1)]
        qb, pb = _i64(data["query"].batch, dev), _i64(data["product"].batch, dev)
        qpos, ppos = _i64(data["query"].pos_emb_id, dev), _i64(data["product"].pos_emb_id, dev)
        cnt = _i64(data["product"].cnt, dev)
        n_graphs = getattr(data, "num_graphs", None)   # (PyG batches and SessionBatch both carry it: no device read)
        if n_graphs is None:
            n_graphs = int(max(int(qb.max()), int(pb.max()))) + 1
        n_graphs = int(n_graphs)
        keep = [xq, xp, qb, pb, qpos, ppos, cnt] + rows
        b = _lib.GraphBatch(n_graphs, qb.shape[0], pb.shape[0], ppos.shape[0],
                            xq.data_ptr() if xq is not None else None, xp.data_ptr() if xp is not None else None,
                            qb.data_ptr(), pb.data_ptr(), qpos.data_ptr(), cnt.data_ptr(), ppos.data_ptr(),
                            rows[0].shape[0], rows[0].data_ptr(), rows[1].data_ptr(),
                            rows[2].shape[0], rows[2].data_ptr(), rows[3].data_ptr(),
                            rows[4].shape[0], rows[4].data_ptr(), rows[5].data_ptr())
        return b, keep, n_graphs

    def _run(self, b, out=None, zq=None, zp=None, run_gnn=True, run_pooling=True, defer_check=False):
        dev = torch.device("cuda", self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        io = _lib.EncoderIO(out.data_ptr() if out is not None else None, zq.data_ptr() if zq is not None else None,
                            zp.data_ptr() if zp is not None else None, int(run_gnn), int(run_pooling), flag.data_ptr())
        check(self._lib.sss_encoder_forward_ex(self._h, ctypes.byref(b), ctypes.byref(io),
                                               _lib.current_stream(self.device)))
        if defer_check:                       # the caller collects the flags and checks them once (check_flags)
            self._pending.append(flag)
            return
        if int(flag.item()) != 0:             # the reference's isnan asserts (model/model.py:301-314)
            raise RuntimeError("nan in embedding[query]")

    def check_flags(self):
        """the NaN checks of every forward run with defer_check=True since the last call, in ONE device read-back"""
        if self._pending:
            bad = int(torch.stack(self._pending).sum().item())
            self._pending = []
            if bad != 0:
                raise RuntimeError("nan in embedding[query]")

    @property
    def node_dim(self):
        return self.in_dim + self.n_layers * self.hidden

    def __call__(self, data, query_node_mask=None, product_node_mask=None, get_node=False, get_token=False,
                 defer_check=False):
        """encoder(data, ...) of model/model.py:279-351.  defer_check=True skips the per-call device read-back of the
        NaN flag (the reference's three isnan asserts cost it three host syncs per batch); call check_flags() after a
        run of batches instead — the host can then prepare batch i + 1 while the GPU encodes batch i.  get_node=True also returns the node embeddings
        {'query': [N_q, 3168], 'product': [N_p, 3168]}; get_token=True returns the reference's (empty)
        session_level_token_emb dict — the block that would fill it is commented out there (model/model.py:321-332)."""
        dev = torch.device("cuda", self.device)
        xq = _f32(data["query"].x, dev)
        xp = _f32(data["product"].input_ids, dev)
        if query_node_mask is not None:       # model/model.py:293-296
            xq = xq * query_node_mask.to(dev).view(-1, 1)
        if product_node_mask is not None:
            xp = xp * product_node_mask.to(dev).view(-1, 1)
        b, keep, n_graphs = self._batch(data, xq, xp)
        out = torch.empty((n_graphs, self.out_dim), dtype=torch.float32, device=dev)
        zq = zp = None
        if get_node:
            zq = torch.empty((xq.shape[0], self.node_dim), dtype=torch.float32, device=dev)
            zp = torch.empty((xp.shape[0], self.node_dim), dtype=torch.float32, device=dev)
        self._run(b, out, zq, zp, defer_check=defer_check)
        self._keep = keep if defer_check else None   # (inputs of an unsynchronised forward must outlive the call)
        del keep
        if not get_node and not get_token:
            return out
        if get_node and not get_token:
            return out, {"query": zq, "product": zp}
        if get_token and not get_node:
            return out, {}
        return out, {"query": zq, "product": zp}, {}

    def gnn(self, x_dict, edge_index_dict, edge_weight_dict=None, add_input_feat=True, data=None):
        """gnn(x_dict, edge_index_dict, edge_weight_dict=None, add_input_feat=True) of model/gnn.py:64-81: the three
        HeteroConv layers; returns {'query': [N_q, in + 3 * hidden], 'product': ...} (without the input block when
        add_input_feat is False)."""
        if edge_weight_dict is not None:
            raise NotImplementedError("the reference never passes edge weights on this path (model/model.py:317); "
                                      "PyG 2.0.4's GATConv would take them as its `size` argument")
        dev = torch.device("cuda", self.device)
        xq, xp = _f32(x_dict["query"], dev), _f32(x_dict["product"], dev)
        nq, np_ = xq.shape[0], xp.shape[0]
        z64 = torch.zeros(1, dtype=torch.int64, device=dev)
        rows = []
        for key in (EDGE_QP, EDGE_PQ, EDGE_PP):
            e = _i64(edge_index_dict[key], dev)
            rows += [e[0].contiguous(), e[1].contiguous()]
        # the GNN stage reads neither batch vectors nor positions: one graph, placeholders of the right length
        qb = torch.zeros(nq, dtype=torch.int64, device=dev)
        pb = torch.zeros(np_, dtype=torch.int64, device=dev)
        cnt = torch.ones(np_, dtype=torch.int64, device=dev)
        b = _lib.GraphBatch(1, nq, np_, np_, xq.data_ptr(), xp.data_ptr(), qb.data_ptr(), pb.data_ptr(), qb.data_ptr(),
                            cnt.data_ptr(), pb.data_ptr(),
                            rows[0].shape[0], rows[0].data_ptr(), rows[1].data_ptr(),
                            rows[2].shape[0], rows[2].data_ptr(), rows[3].data_ptr(),
                            rows[4].shape[0], rows[4].data_ptr(), rows[5].data_ptr())
        zq = torch.empty((nq, self.node_dim), dtype=torch.float32, device=dev)
        zp = torch.empty((np_, self.node_dim), dtype=torch.float32, device=dev)
        self._run(b, None, zq, zp, run_gnn=True, run_pooling=False)
        del z64
        if add_input_feat:
            return {"query": zq, "product": zp}
        return {"query": zq[:, self.in_dim:].contiguous(), "product": zp[:, self.in_dim:].contiguous()}

    def pooling(self, node_emb_dict, data):
        """pooling(input_emb, data) of model/gnn.py:193-217: node embeddings [N, in + 3 * hidden] -> [B, out_dim]"""
        dev = torch.device("cuda", self.device)
        zq, zp = _f32(node_emb_dict["query"], dev), _f32(node_emb_dict["product"], dev)
        if zq.shape[1] != self.node_dim or zp.shape[1] != self.node_dim:
            raise ValueError("pooling expects node embeddings of width %d" % self.node_dim)
        b, keep, n_graphs = self._batch(data, None, None)
        out = torch.empty((n_graphs, self.out_dim), dtype=torch.float32, device=dev)
        self._run(b, out, zq, zp, run_gnn=False, run_pooling=True)
        del keep
        return out


def masked_mean_pool(token_emb, attention_mask, get_token=False):
    """The tail of PretrainedQAEAEncoder.__call__ (model/NodeEmbedding.py:112-125, lin=None): the transformer's
    last_hidden_state [N, L, H] and the tokenizer's attention_mask [N, L] -> sum_t tok * mask / sum_t mask, detached;
    with get_token=True also the token embeddings, as the reference returns them."""
    lib = _lib.load()
    if token_emb.device.type != "cuda":
        raise RuntimeError("masked_mean_pool runs on a CUDA device (no CPU fallback)")
    dev = token_emb.device
    tok = token_emb.detach().to(torch.float32).contiguous()
    mask = attention_mask.to(device=dev, dtype=torch.int64).contiguous()
    n, L, H = tok.shape
    out = torch.empty((n, H), dtype=torch.float32, device=dev)
    check(lib.sss_masked_mean(tok.data_ptr(), mask.data_ptr(), n, L, H, out.data_ptr(), dev.index,
                              _lib.current_stream(dev.index)))
    return (out, token_emb) if get_token else out


def cosine_matrix(a, b):
    """F.normalize(a) @ F.normalize(b).T (fine_tune_ours.py:133,480,494,613,626) on CUDA tensors [na, d], [nb, d]"""
    lib = _lib.load()
    if a.device.type != "cuda":
        raise RuntimeError("cosine_matrix runs on a CUDA device (no CPU fallback)")
    dev = a.device
    a = a.detach().to(torch.float32).contiguous()
    b = b.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=dev)
    check(lib.sss_cosine_matrix(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], a.shape[1], out.data_ptr(),
                                dev.index, _lib.current_stream(dev.index)))
    return out


class BinarizeHead:
    """eval forward of BinarizeHead(n_input, n_output, None) (model/model.py:105-138): sign(lin1(x))"""

    def __init__(self, weight, bias, device=None):
        self._lib = _lib.load()
        self.device = _lib.current_device() if device is None else int(device)
        dev = torch.device("cuda", self.device)
        self.weight, self.bias = _f32(weight, dev), _f32(bias, dev)

    def __call__(self, x):
        dev = torch.device("cuda", self.device)
        x = _f32(x, dev)
        out = torch.empty((x.shape[0], self.weight.shape[0]), dtype=torch.float32, device=dev)
        check(self._lib.sss_binarize_head(x.data_ptr(), self.weight.data_ptr(), self.bias.data_ptr(), x.shape[0],
                                          self.weight.shape[1], self.weight.shape[0], out.data_ptr(), self.device,
                                          _lib.current_stream(self.device)))
        return out


def gather_rows(table, ids):
    """table[ids] on the device through sss_gather_rows (CUDA tensors: fp32 [n_rows, d], int64 [n])"""
    lib = _lib.load()
    dev = table.device
    if dev.type != "cuda":
        raise RuntimeError("gather_rows runs on a CUDA device (no CPU fallback)")
    table = table.contiguous()
    ids = ids.to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty((ids.shape[0], table.shape[1]), dtype=torch.float32, device=dev)
    check(lib.sss_gather_rows(table.data_ptr(), table.shape[0], table.shape[1], ids.data_ptr(), ids.shape[0],
                              out.data_ptr(), dev.index, _lib.current_stream(dev.index)))
    return out


class NodeAsinEmbedding:
    """the reference's per-product id embedding (model/NodeEmbedding.py:128-138): forward(ids) = weight[ids].
    Its output is discarded on the configured path (use_id_embedding=False, model/model.py:288-291); kept for
    drop-in completeness."""

    def __init__(self, weight, device=None):
        self.device = _lib.current_device() if device is None else int(device)
        self.weight = _f32(weight, torch.device("cuda", self.device))

    def __call__(self, ids):
        return gather_rows(self.weight, torch.as_tensor(ids))
