"""HeteroData / Batch with the PyG 2.0.4 collate rules the reference relies on
(util_amazon_filtered.py:98-230 builds the graphs; test_amazon_filterd.py:488 batches them):
  * tensor attributes are concatenated along dim 0, `edge_index` along dim -1;
  * `edge_index` of edge type (src, rel, dst) is offset by the cumulative (src, dst) node counts;
  * every node type gets a `.batch` vector (graph id per node) and `.ptr`;
  * non-tensor attributes are collected into lists; None attributes stay None;
  * edge_index_dict / x_dict iterate in insertion order of the first graph."""
import torch


class _Store(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def to(self, device):
        for k, v in list(self.items()):
            if isinstance(v, torch.Tensor):
                self[k] = v.to(device)
        return self


class HeteroData:
    def __init__(self):
        object.__setattr__(self, "_stores", {})

    def _key(self, k):
        return tuple(k) if isinstance(k, (tuple, list)) else k

    def __getitem__(self, k):
        k = self._key(k)
        st = self._stores
        if k not in st:
            st[k] = _Store()
        return st[k]

    def __setitem__(self, k, v):
        # the reference does data['ori_seq'] = (seq, tar): a graph-level attribute stored as a store
        self._stores[self._key(k)] = v

    @property
    def node_types(self):
        return [k for k, v in self._stores.items() if isinstance(k, str) and isinstance(v, _Store)
                and ("x" in v or "num_nodes" in v)]

    @property
    def edge_types(self):
        return [k for k in self._stores if isinstance(k, tuple)]

    def metadata(self):
        return self.node_types, self.edge_types

    @property
    def edge_index_dict(self):
        return {k: v["edge_index"] for k, v in self._stores.items() if isinstance(k, tuple) and "edge_index" in v}

    @property
    def x_dict(self):
        return {k: v["x"] for k, v in self._stores.items() if isinstance(k, str) and isinstance(v, _Store) and "x" in v}

    def num_nodes_of(self, t):
        s = self._stores[t]
        if "num_nodes" in s:
            return int(s["num_nodes"])
        if "x" in s:
            return int(s["x"].size(0))
        raise ValueError("cannot infer num_nodes of " + str(t))

    def to(self, device):
        for v in self._stores.values():
            if isinstance(v, _Store):
                v.to(device)
        return self


class Batch(HeteroData):
    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        first = data_list[0]
        object.__setattr__(out, "num_graphs", len(data_list))
        counts = {}
        for key, st0 in first._stores.items():
            if not isinstance(st0, _Store):
                out._stores[key] = [d._stores[key] for d in data_list]
                continue
            st = out[key]
            is_node = isinstance(key, str) and ("x" in st0 or "num_nodes" in st0)
            if is_node:
                n = [d.num_nodes_of(key) for d in data_list]
                counts[key] = n
                st["batch"] = torch.repeat_interleave(torch.arange(len(n)), torch.tensor(n))
                st["ptr"] = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(torch.tensor(n), 0)])
            for attr, v0 in st0.items():
                vals = [d._stores[key][attr] for d in data_list]
                if attr == "num_nodes":
                    st[attr] = int(sum(int(v) for v in vals))
                elif v0 is None:
                    st[attr] = None
                elif isinstance(v0, torch.Tensor) and v0.dim() > 0:
                    if attr == "edge_index":
                        src, _, dst = key
                        ns = [d.num_nodes_of(src) for d in data_list]
                        nd = [d.num_nodes_of(dst) for d in data_list]
                        off_s, off_d, parts = 0, 0, []
                        for v, a, b in zip(vals, ns, nd):
                            parts.append(v + torch.tensor([[off_s], [off_d]], dtype=v.dtype))
                            off_s += a
                            off_d += b
                        st[attr] = torch.cat(parts, dim=-1)
                    else:
                        st[attr] = torch.cat(vals, dim=0)
                elif isinstance(v0, torch.Tensor):
                    st[attr] = torch.stack(vals)
                elif isinstance(v0, (int, float)):
                    st[attr] = torch.tensor(vals)
                else:
                    st[attr] = vals
        return out


Data = HeteroData


class Dataset(torch.utils.data.Dataset):
    pass


class InMemoryDataset(Dataset):
    pass
