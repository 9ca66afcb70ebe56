"""Generate golden vectors from the reference's OWN code, run in the build container.

    python tests/golden/gen_golden.py      (needs /root/reference; not needed at test time)

The reference cannot be imported as a package here (torch_geometric / faiss / Levenshtein / tkinter are
absent), so:
  * util_amazon_filtered.normalize is obtained by importing the reference module with placeholder modules
    registered for its two missing imports (torch_geometric.data, Levenshtein) — the function itself is pure
    numpy and runs unmodified;
  * fine_tune_ours.normalize is obtained by exec'ing the three source lines of that function (read from
    the reference file at generation time, never stored in this repo) in a namespace holding numpy.
Outputs: tests/golden/normalize_golden.npz
"""
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "normalize_golden.npz")


def reference_util_normalize():
    tg = types.ModuleType("torch_geometric")
    tgd = types.ModuleType("torch_geometric.data")
    tgd.HeteroData = type("HeteroData", (), {})
    tg.data = tgd
    sys.modules.setdefault("torch_geometric", tg)
    sys.modules.setdefault("torch_geometric.data", tgd)
    sys.modules.setdefault("Levenshtein", types.ModuleType("Levenshtein"))
    sys.path.insert(0, REF)
    mod = importlib.import_module("util_amazon_filtered")
    return mod.normalize


def reference_ft_normalize():
    lines = open(os.path.join(REF, "fine_tune_ours.py")).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def normalize("))
    end = start + 1
    while end < len(lines) and (lines[end].startswith((" ", "\t")) and lines[end].strip()):
        end += 1
    ns = {"np": np}
    exec("\n".join(lines[start:end]), ns)
    return ns["normalize"]


def main():
    rng = np.random.default_rng(20261018)
    util_norm = reference_util_normalize()
    ft_norm = reference_ft_normalize()
    cases = {}
    for name, shape in [("d128", (64, 128)), ("d200", (16, 200)), ("d1600", (8, 1600)), ("d7", (5, 7))]:
        x = rng.standard_normal(shape).astype(np.float32)
        x[0] *= 1e-5   # exercises the 1e-6 clip / +1e-4 floor
        x[1] = 0.0     # zero row
        cases["in_" + name] = x
        cases["util_" + name] = util_norm(x).astype(np.float32)
        cases["ft_" + name] = ft_norm(x).astype(np.float32)
    ones = np.ones(4)
    cases["util_ones4"] = util_norm(ones)   # the reference's only self-check: test_amazon_filterd.py:866
    v = rng.standard_normal(300).astype(np.float32)
    cases["in_vec300"] = v
    cases["util_vec300"] = util_norm(v).astype(np.float32)
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, {k: v.shape for k, v in cases.items()})


if __name__ == "__main__":
    main()
