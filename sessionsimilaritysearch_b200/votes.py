"""Item vote after the neighbour search: get_prediction_by_knn of test_amazon_filterd.py:59-78 on the GPU
(sss_item_vote, csrc/item_vote.cu)."""
import numpy as np
import torch

from . import _lib
from ._lib import check


class ItemLists:
    """items of every database session as a CSR pair on the device (built once per dataset)"""

    def __init__(self, dataset, device=None):
        self.device = _lib.current_device() if device is None else int(device)
        lens, flat = [], []
        for g in dataset:
            x = g['product'].x if not isinstance(g, (list, tuple, np.ndarray)) else g
            x = np.asarray(x.cpu() if torch.is_tensor(x) else x, dtype=np.int64).ravel()
            lens.append(len(x))
            flat.append(x)
        dev = torch.device("cuda", self.device)
        self.n = len(lens)
        self.max_len = int(max(lens)) if lens else 0
        self.item_off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
        self.items = torch.from_numpy(np.concatenate(flat) if flat else np.zeros(0, np.int64)).to(dev)


def item_vote(D, I, lists, K):
    """D [nq, s] neighbour similarities, I [nq, s] neighbour session ids -> (items int64 [nq, K], weights [nq, K])"""
    lib = _lib.load()
    dev = torch.device("cuda", lists.device)
    Dt = (D if torch.is_tensor(D) else torch.from_numpy(np.ascontiguousarray(D, np.float32))).to(dev, torch.float32).contiguous()
    It = (I if torch.is_tensor(I) else torch.from_numpy(np.ascontiguousarray(I, np.int64))).to(dev, torch.int64).contiguous()
    nq, s = Dt.shape
    out_i = torch.empty((nq, K), dtype=torch.int64, device=dev)
    out_w = torch.empty((nq, K), dtype=torch.float32, device=dev)
    check(lib.sss_item_vote(Dt.data_ptr(), It.data_ptr(), nq, s, lists.item_off.data_ptr(), lists.items.data_ptr(),
                            lists.n, K, s * lists.max_len, out_i.data_ptr(), out_w.data_ptr(), lists.device,
                            _lib.current_stream(lists.device)))
    return out_i, out_w


_cache = {}


def get_prediction_by_knn(emb, index, dataset, sample_size, K):
    """same signature and result as the reference: ids of the K items with the largest summed neighbour similarity
    (float64 sums in arrival order, equal sums in order of first arrival — pinned by tests/golden/vote_golden.npz)"""
    key = id(dataset)
    if key not in _cache:
        _cache.clear()
        _cache[key] = ItemLists(dataset, device=getattr(index, "device", None))
    q = emb.detach() if torch.is_tensor(emb) else emb
    if torch.is_tensor(q) and q.dim() == 1:
        q = q.view(1, -1)
    D, I = index.search(q if (torch.is_tensor(q) and q.is_cuda) else np.asarray(q.cpu() if torch.is_tensor(q) else q), sample_size)
    items, _ = item_vote(D, I, _cache[key], K)
    row = items[0].tolist()
    return [i for i in row if i >= 0]
