"""Golden vectors of the neighbour item vote, produced by the reference's OWN get_prediction_by_knn.

    python tests/golden/gen_vote_golden.py      (needs /root/reference; not needed at test time)

test_amazon_filterd.py cannot be imported here (faiss / torch_geometric are absent), so the source lines of
`get_prediction_by_knn` (test_amazon_filterd.py:59-78) are read from the reference file at generation time — never
stored in this repo — and exec'd unmodified in a namespace holding numpy and collections.defaultdict.  The function
only needs `index.search(emb, sample_size) -> (D, I)` and `dataset[i]['product'].x`, which are supplied by stand-ins
that replay seeded (D, I) and item lists.  What gets pinned: float64 accumulation in arrival order and the stable sort
(equal weights keep first-arrival order), including heavy ties and the reference's own call shape
(sample_size = 500, K = 20, test_amazon_filterd.py:189-199).
Output: tests/golden/vote_golden.npz
"""
import os
from collections import defaultdict

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vote_golden.npz")


def reference_fn():
    lines = open(os.path.join(REF, "test_amazon_filterd.py")).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def get_prediction_by_knn("))
    end = start + 1
    while end < len(lines) and (not lines[end].strip() or lines[end].startswith((" ", "\t"))):
        end += 1
    ns = {"np": np, "defaultdict": defaultdict}
    exec("\n".join(lines[start:end]), ns)
    return ns["get_prediction_by_knn"]


class ReplayIndex:
    def __init__(self, D, I):
        self.D, self.I = D, I

    def search(self, emb, s):
        assert s == self.D.shape[1]
        return self.D.copy(), self.I.copy()


class Node:
    def __init__(self, x):
        self.x = x


def make_case(rng, n_sessions, s, max_items, n_items, tie_mode):
    lens = rng.integers(1, max_items + 1, size=n_sessions)
    item_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    items = np.concatenate([rng.choice(n_items, size=l, replace=False) + 1 for l in lens]).astype(np.int64)
    I = rng.choice(n_sessions, size=s, replace=False).astype(np.int64)
    if tie_mode == "heavy":      # a handful of distinct similarities: equal sums everywhere
        D = rng.choice(np.array([0.5, 0.25, 0.75], np.float32), size=s)
    elif tie_mode == "equal":    # every neighbour has the same weight
        D = np.full(s, 0.625, np.float32)
    else:                        # cosine-like, descending as a search returns them
        D = np.sort(rng.uniform(0.2, 0.99, size=s).astype(np.float32))[::-1].copy()
    return item_off, items, D.reshape(1, s).astype(np.float32), I.reshape(1, s)


def main():
    fn = reference_fn()
    rng = np.random.default_rng(20261019)
    out = {}
    cases = [("small", 50, 8, 6, 40, "cos", 5), ("heavy_ties", 300, 60, 12, 80, "heavy", 20),
             ("equal_weights", 300, 40, 10, 60, "equal", 20), ("ref_call_500", 4000, 500, 19, 3000, "cos", 20),
             ("ref_call_500_ties", 4000, 500, 19, 200000, "heavy", 20)]
    for name, n_sessions, s, max_items, n_items, tie_mode, K in cases:
        item_off, items, D, I = make_case(rng, n_sessions, s, max_items, n_items, tie_mode)
        dataset = [{"product": Node(items[item_off[i]:item_off[i + 1]])} for i in range(n_sessions)]
        pred = fn(torch.zeros(1, 4), ReplayIndex(D, I), dataset, s, K)
        exp = np.full(K, -1, np.int64)
        exp[:len(pred)] = np.asarray(pred, np.int64)
        out[name + "_item_off"], out[name + "_items"] = item_off, items
        out[name + "_D"], out[name + "_I"], out[name + "_K"], out[name + "_expected"] = D, I, np.int64(K), exp
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.endswith("_expected")})


if __name__ == "__main__":
    main()
