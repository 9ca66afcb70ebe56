"""Golden vectors of the evaluation metrics, produced by the reference's OWN get_score (fine_tune_ours.py:42-87).

    python tests/golden/gen_metrics_golden.py      (needs /root/reference; not needed at test time)

fine_tune_ours.py cannot be imported (tkinter / faiss / torch_geometric / a missing in-repo module), so the source
lines of `get_score` are read from the reference file at generation time — never stored in this repo — and exec'd
unmodified in a namespace holding numpy and the reference's own helpers (`get_item`, `get_item_type`, `get_query`,
`get_session_item_title`, imported from /root/reference/util_amazon_filtered.py with placeholder modules for its two
missing imports).  `Levenshtein` is absent here, so only the three sim types that do not touch it are pinned:
all_jaccard, cur_jaccard, all_product_type_score.
Sessions come from sessionsimilaritysearch_b200.synth.make_sessions (seeded); the fixture stores the seeds and the
scores, the tests rebuild the sessions.
Output: tests/golden/metrics_golden.npz
"""
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
OUT = os.path.join(HERE, "metrics_golden.npz")


def reference_get_score():
    tg = types.ModuleType("torch_geometric")
    tgd = types.ModuleType("torch_geometric.data")
    tgd.HeteroData = type("HeteroData", (), {})
    tg.data = tgd
    sys.modules.setdefault("torch_geometric", tg)
    sys.modules.setdefault("torch_geometric.data", tgd)
    sys.modules.setdefault("Levenshtein", types.ModuleType("Levenshtein"))
    sys.path.insert(0, REF)
    util = importlib.import_module("util_amazon_filtered")
    lines = open(os.path.join(REF, "fine_tune_ours.py")).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def get_score("))
    end = start + 1
    while end < len(lines) and (not lines[end].strip() or lines[end].startswith((" ", "\t"))):
        end += 1
    ns = {"np": np, "get_item": util.get_item, "get_item_type": util.get_item_type, "get_query": util.get_query,
          "get_session_item_title": util.get_session_item_title, "Levenshtein": sys.modules["Levenshtein"]}
    exec("\n".join(lines[start:end]), ns)
    return ns["get_score"]


def make_data(seed_test, seed_train, n_test, n_train, n_items=None):
    from sessionsimilaritysearch_b200 import synth
    rng = np.random.default_rng(seed_test + 7)
    test = [synth.split_session(s, rng) for s in synth.make_sessions(n_test, seed_test)]
    train = synth.make_sessions(n_train, seed_train)
    return test, train


def main():
    get_score = reference_get_score()
    test, train = make_data(101, 202, 40, 300)
    rng = np.random.default_rng(303)
    I = rng.integers(0, len(train), size=(len(test), 25)).astype(np.int64)
    out = {"seed_test": np.int64(101), "seed_train": np.int64(202), "n_test": np.int64(40), "n_train": np.int64(300),
           "I": I}
    for sim in ("all_jaccard", "cur_jaccard", "all_product_type_score"):
        gt = np.zeros_like(I, dtype=np.float32)
        for i, t in enumerate(test):
            for j, d in enumerate(I[i, :]):
                gt[i, j] = get_score(t, (train[d], []), sim)   # the loop of fine_tune_ours.py:883-888
        out["gt_" + sim] = gt
        out["mean_" + sim] = np.float32(np.mean(gt))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: float(v) for k, v in out.items() if k.startswith("mean_")})


if __name__ == "__main__":
    main()
