import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import search_oracle
    search_oracle.build()
    return search_oracle


# ---- synthetic data -----------------------------------------------------------------

def make_iid(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def make_clustered(n, d, seed, n_centroids=1024, noise=0.3):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((n_centroids, d)).astype(np.float32)
    a = rng.integers(0, n_centroids, size=n)
    return (c[a] + noise * rng.standard_normal((n, d))).astype(np.float32)


def make_ties(n, d, seed):
    """entries in {-1,+1}: integer scores, ties everywhere -> exercises the (score desc, id asc) rule"""
    return (np.random.default_rng(seed).integers(0, 2, size=(n, d)) * 2 - 1).astype(np.float32)


def make_segments(n_rows, seed, mean=7):
    """sessions with 1+Poisson(mean) contiguous subsession rows until n_rows are used"""
    rng = np.random.default_rng(seed)
    lens = []
    total = 0
    while total < n_rows:
        l = int(1 + rng.poisson(mean))
        l = min(l, n_rows - total)
        lens.append(l)
        total += l
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)


def make_session_rows(seg_off, d, seed, noise=0.3):
    """rows of one session are correlated: base_i + noise * eps_ij (prefix embeddings of one session)"""
    rng = np.random.default_rng(seed)
    n_seg = len(seg_off) - 1
    base = rng.standard_normal((n_seg, d)).astype(np.float32)
    rep = np.repeat(np.arange(n_seg), np.diff(seg_off))
    return (base[rep] + noise * rng.standard_normal((len(rep), d))).astype(np.float32)
