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
        self.load_state_dict(params)
        # dense linears: this library's split-bf16 tcgen05 GEMM by default (2.7e-5 of the output scale from float64);
        # math="fp32" selects cuBLAS' pedantic sgemm (2.5e-6), the C ABI's own default
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
        """'fp32' (cuBLAS pedantic sgemm, default), 'bf16x3' (this library's tcgen05 GEMM on split-bf16 operands:
        fp32-level accuracy on the tensor cores) or 'bf16x9' (cuBLAS' fp32 emulation; raises RuntimeError when the
        loaded cuBLAS does not offer it)"""
        check(self._lib.sss_encoder_set_math(self._h, {"fp32": 0, "bf16x9": 1, "bf16x3": 2}[math]))
        return self

    @property
    def math(self):
        return ("fp32", "bf16x9", "bf16x3")[int(self._lib.sss_encoder_get_math(self._h))]

    def eval(self):
        return self

    def to(self, device):
        return self

    def __call__(self, data, query_node_mask=None, product_node_mask=None, get_node=False, get_token=False):
        if get_node or get_token:
            raise NotImplementedError("node / token level outputs are training-only paths of the reference")
        dev = torch.device("cuda", self.device)
        xq = _f32(data["query"].x, dev)
        xp = _f32(data["product"].input_ids, dev)
        if query_node_mask is not None:       # model/model.py:293-296
            xq = xq * query_node_mask.to(dev).view(-1, 1)
        if product_node_mask is not None:
            xp = xp * product_node_mask.to(dev).view(-1, 1)
        ei = data.edge_index_dict
        qp, pq, pp = _i64(ei[EDGE_QP], dev), _i64(ei[EDGE_PQ], dev), _i64(ei[EDGE_PP], dev)
        qb, pb = _i64(data["query"].batch, dev), _i64(data["product"].batch, dev)
        qpos, ppos = _i64(data["query"].pos_emb_id, dev), _i64(data["product"].pos_emb_id, dev)
        cnt = _i64(data["product"].cnt, dev)
        n_graphs = int(max(int(qb.max()), int(pb.max()))) + 1
        keep = [xq, xp, qp, pq, pp, qb, pb, qpos, ppos, cnt]
        b = _lib.GraphBatch(n_graphs, xq.shape[0], xp.shape[0], ppos.shape[0], xq.data_ptr(), xp.data_ptr(),
                            qb.data_ptr(), pb.data_ptr(), qpos.data_ptr(), cnt.data_ptr(), ppos.data_ptr(),
                            qp.shape[1], qp[0].contiguous().data_ptr(), qp[1].contiguous().data_ptr(),
                            pq.shape[1], pq[0].contiguous().data_ptr(), pq[1].contiguous().data_ptr(),
                            pp.shape[1], pp[0].contiguous().data_ptr(), pp[1].contiguous().data_ptr())
        out = torch.empty((n_graphs, self.out_dim), dtype=torch.float32, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        check(self._lib.sss_encoder_forward(self._h, ctypes.byref(b), out.data_ptr(), flag.data_ptr(),
                                            _lib.current_stream(self.device)))
        del keep
        if int(flag.item()) != 0:             # the reference's isnan asserts (model/model.py:301-314)
            raise RuntimeError("nan in embedding[query]")
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
