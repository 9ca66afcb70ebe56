"""Session graphs and their batches: the container the reference gets from PyG's HeteroData / Batch /
DataLoader (util_amazon_filtered.py:99; test_amazon_filterd.py:488,547; fine_tune_ours.py:790,814), rebuilt as
a small purpose-made structure: one attribute bag per node type / edge type, and a batcher that concatenates
node attributes, offsets each edge type's `edge_index` by the (source, destination) node counts seen so far,
and adds `.batch` / `.ptr` per node type.  Only what the encoder path reads is modelled.
"""
import torch

EDGE_QP = ("query", "clicks", "product")
EDGE_PQ = ("product", "clicked by", "query")
EDGE_PP = ("product", "to", "product")


class AttrBag:
    """attribute bag with dict and dotted access (data['query'].x  /  data['query']['x'])"""

    def __init__(self):
        self.__dict__["_a"] = {}

    def __getattr__(self, name):
        a = self.__dict__["_a"]
        if name in a:
            return a[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self.__dict__["_a"][name] = value

    __getitem__ = __getattr__
    __setitem__ = __setattr__

    def __contains__(self, name):
        return name in self.__dict__["_a"]

    def keys(self):
        return self.__dict__["_a"].keys()

    def items(self):
        return self.__dict__["_a"].items()

    def to(self, device, non_blocking=False):
        a = self.__dict__["_a"]
        for k, v in a.items():
            if torch.is_tensor(v):
                a[k] = v.to(device, non_blocking=non_blocking)
        return self


class SessionGraph:
    """one session as a heterogeneous graph (node types 'query', 'product', ...; three edge types)"""

    def __init__(self):
        self.bags = {}     # insertion ordered: node types (str), edge types (3-tuples), graph attributes
        self.extras = {}   # plain graph-level attributes such as 'ori_seq'

    def __getitem__(self, key):
        key = tuple(key) if isinstance(key, list) else key
        if key in self.extras:
            return self.extras[key]
        if key not in self.bags:
            self.bags[key] = AttrBag()
        return self.bags[key]

    def __setitem__(self, key, value):
        self.extras[key] = value

    @property
    def node_types(self):
        return [k for k, b in self.bags.items() if isinstance(k, str) and ("x" in b or "num_nodes" in b)]

    @property
    def edge_types(self):
        return [k for k in self.bags if isinstance(k, tuple)]

    def metadata(self):
        return self.node_types, self.edge_types

    @property
    def edge_index_dict(self):
        return {k: b.edge_index for k, b in self.bags.items() if isinstance(k, tuple) and "edge_index" in b}

    def node_count(self, t):
        b = self.bags[t]
        return int(b.num_nodes) if "num_nodes" in b else int(b.x.shape[0])

    def to(self, device, non_blocking=False):
        for b in self.bags.values():
            b.to(device, non_blocking)
        return self


class SessionBatch(SessionGraph):
    """several SessionGraphs glued into one disconnected graph"""
    num_graphs = 0


def collate(graphs):
    out = SessionBatch()
    out.num_graphs = len(graphs)
    g0 = graphs[0]
    sizes = {t: torch.tensor([g.node_count(t) for g in graphs], dtype=torch.long) for t in g0.node_types}
    starts = {t: torch.cumsum(n, 0) - n for t, n in sizes.items()}
    for key, bag0 in g0.bags.items():
        dst = out[key]
        if key in sizes:
            n = sizes[key]
            dst.batch = torch.repeat_interleave(torch.arange(len(graphs)), n)
            dst.ptr = torch.cat([n.new_zeros(1), torch.cumsum(n, 0)])
        for name, v0 in bag0.items():
            col = [g.bags[key][name] for g in graphs]
            if name == "num_nodes":
                dst[name] = int(sum(int(c) for c in col))
            elif v0 is None:
                dst[name] = None
            elif torch.is_tensor(v0) and v0.dim() == 0:
                dst[name] = torch.stack(col)
            elif torch.is_tensor(v0) and name == "edge_index":
                s, _, d = key
                shift = torch.stack([starts[s], starts[d]], 0)                      # [2, G]
                per_graph = torch.tensor([c.shape[1] for c in col], dtype=torch.long)
                dst[name] = torch.cat(col, 1) + torch.repeat_interleave(shift, per_graph, dim=1)
            elif torch.is_tensor(v0):
                dst[name] = torch.cat(col, 0)
            elif isinstance(v0, (int, float)):
                dst[name] = torch.tensor(col)
            else:
                dst[name] = col
    for key in g0.extras:
        out.extras[key] = [g.extras[key] for g in graphs]
    return out


class DataLoader(torch.utils.data.DataLoader):
    """DataLoader(graph_list, batch_size=200, shuffle=False) as used at test_amazon_filterd.py:488"""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        kw.pop("collate_fn", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, collate_fn=collate, **kw)
