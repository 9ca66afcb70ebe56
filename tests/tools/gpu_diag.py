"""First-contact diagnostic for a B200 box: runs each mode in its own subprocess with a timeout so that one
failing kernel cannot hide the others.  Usage: python scripts/gpu_diag.py"""
import subprocess
import sys
import textwrap

CASES = {
    "normalize": """
        x = rng.standard_normal((1000, 128)).astype(np.float32)
        g = sss.normalize(x); o = so.normalize(x, 1)
        print('bit-exact', np.array_equal(g.view(np.uint32), o.view(np.uint32)), np.abs(g-o).max())
    """,
    "fp32": """
        db = rng.standard_normal((20000, 128)).astype(np.float32); q = rng.standard_normal((33, 128)).astype(np.float32)
        ix = sss.build_index(db, 'ip', mode='fp32'); D, I = ix.search(q, 100)
        Do, Io = so.search_flat(db, q, 100)
        print('ids', (I == Io).mean(), 'scores', np.array_equal(D, Do), ix.stats())
    """,
    "bf16": """
        db = rng.standard_normal((20000, 128)).astype(np.float32); q = rng.standard_normal((33, 128)).astype(np.float32)
        ix = sss.build_index(db, 'ip', mode='bf16'); D, I = ix.search(q, 100)
        Do, Io = so.search_flat(db, q, 100)
        rec = np.mean([len(set(I[r]) & set(Io[r])) / 100.0 for r in range(33)])
        print('recall', rec, 'max|dD|', np.abs(D - Do).max(), 'rel', np.abs(D - Do).max() / np.abs(Do).max(), ix.stats())
        print(D[0, :5], Do[0, :5], I[0, :5], Io[0, :5])
    """,
    "exact": """
        db = rng.standard_normal((200000, 128)).astype(np.float32); q = rng.standard_normal((300, 128)).astype(np.float32)
        ix = sss.build_index(db, 'cos', mode='exact'); qn = sss.normalize(q); D, I = ix.search(qn, 100)
        Do, Io = so.search_flat(so.normalize(db, 1), so.normalize(q, 1), 100)
        print('ids', (I == Io).mean(), 'scores', np.array_equal(D, Do), ix.stats())
    """,
}

PRE = """
import sys, numpy as np
sys.path.insert(0, '.')
import sessionsimilaritysearch_b200 as sss
from oracle import search_oracle as so
rng = np.random.default_rng(0)
"""

for name, body in CASES.items():
    code = PRE + textwrap.dedent(body)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240)
        print("== %s rc=%d\n%s%s" % (name, r.returncode, r.stdout[-2000:], r.stderr[-3000:]))
    except subprocess.TimeoutExpired:
        print("== %s TIMEOUT" % name)
    sys.stdout.flush()
