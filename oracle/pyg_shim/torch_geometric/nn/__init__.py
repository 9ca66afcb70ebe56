"""Message-passing layers of torch_geometric.nn 2.0.4 used by model/gnn.py:43-81,183-217, in plain torch.

Recalled semantics (SURVEY.md 8c, to be re-verified against the pinned wheel when one is available):
  GATConv((-1,-1), C), heads=1: bias-free lin_src / lin_dst, additive attention a_s[j] + a_d[i],
      leaky_relu(0.2), softmax over the incoming edges of each destination (exp(e - max) / (sum + 1e-16)),
      sum aggregation, + bias.  add_self_loops=True on bipartite input: remove edges with src == dst
      (batch-global indices), then append (i, i) for i < min(N_src, N_dst).
  GatedGraphConv(C, L): zero-pad x to C, m = x @ weight[l], sum-aggregate m over incoming edges
      (optionally * edge_weight), x = GRUCell(m, x).
  HeteroConv(convs, aggr='sum'): iterate edge_index_dict in order, skip missing convs, tuple input when
      src != dst, per destination type stack(...).sum(0).
  global_{mean,add,max}_pool: segment reduce over `batch`, size = batch.max() + 1.
"""
import math
from collections import defaultdict

import torch
import torch.nn.functional as F
from torch import nn


def _scatter_sum(src, index, dim_size):
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def _scatter_max(src, index, dim_size):
    out = src.new_full((dim_size,) + tuple(src.shape[1:]), float("-inf"))
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    out = out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)  # torch_scatter fills empty with 0


def global_add_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return _scatter_sum(x, batch, size)


def global_mean_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    s = _scatter_sum(x, batch, size)
    cnt = _scatter_sum(torch.ones_like(batch, dtype=x.dtype), batch, size).clamp(min=1)
    return s / cnt.view(-1, *([1] * (x.dim() - 1)))


def global_max_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return _scatter_max(x, batch, size)


def global_sort_pool(*a, **k):
    raise NotImplementedError("global_sort_pool is not on the reference's hot path")


def softmax(src, index, num_nodes):
    src_max = _scatter_max(src, index, num_nodes).index_select(0, index)
    out = (src - src_max).exp()
    out_sum = _scatter_sum(out, index, num_nodes).index_select(0, index)
    return out / (out_sum + 1e-16)


class Linear(nn.Module):
    """torch_geometric.nn.Linear with lazy in_channels (-1) and glorot initialisation"""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer = weight_initializer
        if in_channels > 0:
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
            self._init()
        else:
            self.weight = nn.parameter.UninitializedParameter()
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def _init(self):
        if self.weight_initializer == "glorot":
            a = math.sqrt(6.0 / (self.weight.size(0) + self.weight.size(1)))
            nn.init.uniform_(self.weight, -a, a)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def forward(self, x):
        if isinstance(self.weight, nn.parameter.UninitializedParameter):
            self.in_channels = x.size(-1)
            self.weight.materialize((self.out_channels, self.in_channels))
            self._init()
        return F.linear(x, self.weight, self.bias)


def remove_self_loops(edge_index):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask]


def add_self_loops(edge_index, num_nodes):
    loop = torch.arange(0, num_nodes, dtype=torch.long, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1)


class GATConv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, bias=True):
        super().__init__()
        assert heads == 1, "the reference uses heads=1 (model/gnn.py:54)"
        self.out_channels, self.heads, self.negative_slope = out_channels, heads, negative_slope
        self.add_self_loops = add_self_loops
        if isinstance(in_channels, int):
            self.lin_src = Linear(in_channels, heads * out_channels, bias=False, weight_initializer="glorot")
            self.lin_dst = self.lin_src
        else:
            self.lin_src = Linear(in_channels[0], heads * out_channels, False, weight_initializer="glorot")
            self.lin_dst = Linear(in_channels[1], heads * out_channels, False, weight_initializer="glorot")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels)) if bias else None
        a = math.sqrt(6.0 / (heads + out_channels))
        nn.init.uniform_(self.att_src, -a, a)
        nn.init.uniform_(self.att_dst, -a, a)

    def forward(self, x, edge_index, size=None):
        H, C = self.heads, self.out_channels
        if isinstance(x, torch.Tensor):
            x_src = x_dst = self.lin_src(x).view(-1, H, C)
        else:
            x_src, x_dst = x
            x_src = self.lin_src(x_src).view(-1, H, C)
            if x_dst is not None:
                x_dst = self.lin_dst(x_dst).view(-1, H, C)
        alpha_src = (x_src * self.att_src).sum(dim=-1)
        alpha_dst = None if x_dst is None else (x_dst * self.att_dst).sum(dim=-1)
        n_dst = x_dst.size(0) if x_dst is not None else x_src.size(0)
        if self.add_self_loops:
            num_nodes = x_src.size(0)
            if x_dst is not None:
                num_nodes = min(num_nodes, x_dst.size(0))
            edge_index = remove_self_loops(edge_index)
            edge_index = add_self_loops(edge_index, num_nodes=num_nodes)
        j, i = edge_index[0], edge_index[1]
        alpha = alpha_src.index_select(0, j)
        if alpha_dst is not None:
            alpha = alpha + alpha_dst.index_select(0, i)
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = softmax(alpha, i, n_dst)
        msg = x_src.index_select(0, j) * alpha.unsqueeze(-1)
        out = _scatter_sum(msg, i, n_dst).view(-1, H * C)
        if self.bias is not None:
            out = out + self.bias
        return out


class GatedGraphConv(nn.Module):
    def __init__(self, out_channels, num_layers, aggr="add", bias=True):
        super().__init__()
        self.out_channels, self.num_layers = out_channels, num_layers
        self.weight = nn.Parameter(torch.empty(num_layers, out_channels, out_channels))
        self.rnn = nn.GRUCell(out_channels, out_channels, bias=bias)
        b = 1.0 / math.sqrt(out_channels)
        nn.init.uniform_(self.weight, -b, b)

    def forward(self, x, edge_index, edge_weight=None):
        if x.size(-1) > self.out_channels:
            raise ValueError("The number of input channels is not allowed to be larger than the number of output "
                             "channels")
        if x.size(-1) < self.out_channels:
            x = torch.cat([x, x.new_zeros(x.size(0), self.out_channels - x.size(-1))], dim=1)
        for l in range(self.num_layers):
            m = torch.matmul(x, self.weight[l])
            mj = m.index_select(0, edge_index[0])
            if edge_weight is not None:
                mj = edge_weight.view(-1, 1) * mj
            m = _scatter_sum(mj, edge_index[1], x.size(0))
            x = self.rnn(m, x)
        return x


class HeteroConv(nn.Module):
    def __init__(self, convs, aggr="sum"):
        super().__init__()
        self.convs = nn.ModuleDict({"__".join(k): v for k, v in convs.items()})
        self.aggr = aggr

    def forward(self, x_dict, edge_index_dict, *args_dict):
        out_dict = defaultdict(list)
        for edge_type, edge_index in edge_index_dict.items():
            src, rel, dst = edge_type
            key = "__".join(edge_type)
            if key not in self.convs:
                continue
            args = []
            for value_dict in args_dict:
                if edge_type in value_dict:
                    args.append(value_dict[edge_type])
            conv = self.convs[key]
            if src == dst:
                out = conv(x_dict[src], edge_index, *args)
            else:
                out = conv((x_dict[src], x_dict[dst]), edge_index, *args)
            out_dict[dst].append(out)
        res = {}
        for key, xs in out_dict.items():
            if len(xs) == 1:
                res[key] = xs[0]
            else:
                st = torch.stack(xs, dim=0)
                res[key] = {"sum": st.sum(0), "mean": st.mean(0), "max": st.max(0)[0], "min": st.min(0)[0]}[self.aggr]
        return res


def _absent(name):
    class _Absent(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError(name + " is not on the reference's hot path (SURVEY.md section 2)")
    _Absent.__name__ = name
    return _Absent


GCNConv = _absent("GCNConv")
SAGEConv = _absent("SAGEConv")
HGTConv = _absent("HGTConv")


def to_hetero(*a, **k):
    raise NotImplementedError("to_hetero is not on the reference's hot path")
