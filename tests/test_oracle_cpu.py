"""CPU tests of the oracle itself: golden vectors from the reference's own normalize(), float64 brute
force, tie rule, segments, Hamming, item vote, merge."""
import os

import numpy as np

from conftest import make_clustered, make_iid, make_segments, make_session_rows, make_ties

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "normalize_golden.npz"))


def test_reference_self_check_value():
    # test_amazon_filterd.py:866 prints normalize(np.ones(4)) -> [0.5 0.5 0.5 0.5]
    assert np.allclose(GOLD["util_ones4"], 0.5)


def test_oracle_normalize_matches_reference_golden(oracle):
    for name in ["d128", "d200", "d1600", "d7"]:
        x = GOLD["in_" + name]
        # fixed-order fp32 sum vs numpy pairwise sum: a few ulp
        np.testing.assert_allclose(oracle.normalize(x, oracle.NORM_UTIL), GOLD["util_" + name], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(oracle.normalize(x, oracle.NORM_FT), GOLD["ft_" + name], rtol=2e-6, atol=1e-9)
    v = GOLD["in_vec300"]
    np.testing.assert_allclose(oracle.normalize(v[None, :], oracle.NORM_UTIL)[0], GOLD["util_vec300"], rtol=2e-6)


def test_numpy_restatements_match_reference_golden(oracle):
    for name in ["d128", "d200", "d1600", "d7"]:
        x = GOLD["in_" + name]
        assert np.array_equal(oracle.normalize_util_numpy(x).astype(np.float32), GOLD["util_" + name])
        assert np.array_equal(oracle.normalize_ft_numpy(x).astype(np.float32), GOLD["ft_" + name])


def test_zero_row_normalize(oracle):
    x = np.zeros((2, 16), dtype=np.float32)
    for mode in (oracle.NORM_UTIL, oracle.NORM_FT, oracle.NORM_TORCH):
        assert np.array_equal(oracle.normalize(x, mode), x)


def test_o2_vs_float64(oracle):
    db = oracle.normalize(make_iid(5000, 128, 1), oracle.NORM_UTIL)
    q = oracle.normalize(make_iid(8, 128, 2), oracle.NORM_UTIL)
    D, I = oracle.search_flat(db, q, 20)
    D64, I64 = oracle.search_float64(db, q, 20)
    # ids may differ only inside near-ties of the float64 scores
    for r in range(q.shape[0]):
        diff = np.nonzero(I[r] != I64[r])[0]
        for j in diff:
            assert abs(D64[r, j] - float(D[r, j])) < 1e-6
    np.testing.assert_allclose(D, D64, atol=2e-6)
    assert np.all(np.diff(D, axis=1) <= 0)


def test_o2_l2(oracle):
    db = make_iid(3000, 64, 3)
    q = make_iid(5, 64, 4)
    D, I = oracle.search_flat(db, q, 10, metric=oracle.METRIC_L2)
    ref = ((q[:, None, :].astype(np.float64) - db[None].astype(np.float64)) ** 2).sum(-1)
    idx = np.argsort(ref, axis=1, kind="stable")[:, :10]
    assert np.array_equal(I, idx)
    np.testing.assert_allclose(D, np.take_along_axis(ref, idx, 1), rtol=1e-5)
    assert np.all(np.diff(D, axis=1) >= 0)


def test_tie_rule_and_padding(oracle):
    db = make_ties(500, 32, 5)
    q = make_ties(6, 32, 6)
    D, I = oracle.search_flat(db, q, 50)
    sc = q.astype(np.float64) @ db.T.astype(np.float64)
    for r in range(q.shape[0]):
        order = sorted(range(db.shape[0]), key=lambda i: (-sc[r, i], i))[:50]
        assert list(I[r]) == order
        assert np.array_equal(D[r], sc[r, order].astype(np.float32))
    # fewer rows than k: tail is (-inf, -1)
    D, I = oracle.search_flat(db[:7], q, 10)
    assert np.all(I[:, 7:] == -1) and np.all(np.isneginf(D[:, 7:])) and np.all(I[:, :7] >= 0)
    D, I = oracle.search_flat(db[:7], q, 10, metric=oracle.METRIC_L2)
    assert np.all(I[:, 7:] == -1) and np.all(np.isposinf(D[:, 7:]))


def test_segment_max_and_sum(oracle):
    seg = make_segments(4000, 7)
    db = oracle.normalize(make_session_rows(seg, 48, 8), oracle.NORM_UTIL)
    q = oracle.normalize(make_iid(7, 48, 9), oracle.NORM_UTIL)
    n_seg = len(seg) - 1
    rows = q.astype(np.float64) @ db.T.astype(np.float64)
    smax = np.stack([rows[:, seg[s]:seg[s + 1]].max(1) for s in range(n_seg)], 1)
    ssum = np.stack([rows[:, seg[s]:seg[s + 1]].sum(1) for s in range(n_seg)], 1)
    D, I = oracle.search_flat(db, q, 15, seg_off=seg, reduce=oracle.REDUCE_MAX)
    assert I.max() < n_seg
    np.testing.assert_allclose(D, -np.sort(-smax, axis=1)[:, :15], atol=2e-6)
    D, I = oracle.search_flat(db, q, 15, seg_off=seg, reduce=oracle.REDUCE_SUM)
    np.testing.assert_allclose(D, -np.sort(-ssum, axis=1)[:, :15], atol=1e-5)
    # empty segments never appear
    seg2 = np.array([0, 0, 3, 3, 10], dtype=np.int64)
    D, I = oracle.search_flat(db[:10], q, 4, seg_off=seg2, reduce=oracle.REDUCE_MAX)
    assert set(I[:, :2].ravel()) <= {1, 3} and np.all(I[:, 2:] == -1)


def test_blas_baseline_agrees(oracle):
    db = oracle.normalize(make_clustered(20000, 128, 10), oracle.NORM_UTIL)
    q = oracle.normalize(make_clustered(16, 128, 11), oracle.NORM_UTIL)
    D, I = oracle.search_flat(db, q, 100)
    Db, Ib = oracle.search_blas(db, q, 100, chunk=4096)
    assert (I == Ib).mean() > 0.99
    np.testing.assert_allclose(D, Db, atol=5e-6)
    seg = make_segments(20000, 12)
    D, I = oracle.search_flat(db, q, 50, seg_off=seg, reduce=oracle.REDUCE_MAX)
    Db, Ib = oracle.search_blas(db, q, 50, seg_off=seg, reduce=oracle.REDUCE_MAX, chunk=3000)
    assert (I == Ib).mean() > 0.99


def test_hamming_and_pack(oracle):
    rng = np.random.default_rng(13)
    x = rng.standard_normal((300, 250)).astype(np.float32)
    x[rng.random(x.shape) < 0.05] = 0.0
    s = np.sign(x)                                      # BinarizeHead eval output, model/model.py:137
    ref = np.packbits(((s + 1) / 2).astype(int), axis=1)  # fine_tune_ours.py:839-840
    codes = oracle.pack_sign_bits(s)
    assert np.array_equal(codes, ref) and codes.shape == (300, 32)
    qc = codes[:9].copy()
    qc[:, 3] ^= 0x5A
    D, I = oracle.search_hamming(codes, qc, 12)
    bits = np.unpackbits(codes, axis=1).astype(np.int32)
    qb = np.unpackbits(qc, axis=1).astype(np.int32)
    dist = (qb[:, None, :] != bits[None]).sum(-1)
    for r in range(9):
        order = sorted(range(300), key=lambda i: (dist[r, i], i))[:12]
        assert list(I[r]) == order and list(D[r]) == [dist[r, i] for i in order]


def test_item_vote(oracle):
    rng = np.random.default_rng(14)
    n_sess = 50
    lens = rng.integers(1, 6, size=n_sess)
    item_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    items = rng.integers(0, 40, size=item_off[-1]).astype(np.int64)
    I = np.stack([rng.permutation(n_sess)[:10] for _ in range(4)]).astype(np.int64)
    D = np.sort(rng.random((4, 10)).astype(np.float32), axis=1)[:, ::-1].copy()
    oi, ow = oracle.item_vote(D, I, item_off, items, 5)
    for r in range(4):
        # the reference's defaultdict loop (test_amazon_filterd.py:70-76), restated
        aw = {}
        for j in range(10):
            for it in items[item_off[I[r, j]]:item_off[I[r, j] + 1]]:
                aw[int(it)] = np.float32(aw.get(int(it), np.float32(0)) + D[r, j])
        exp = sorted(aw.items(), key=lambda kv: (-kv[1], kv[0]))[:5]
        assert [e[0] for e in exp] == list(oi[r][:len(exp)])
        np.testing.assert_allclose([e[1] for e in exp], ow[r][:len(exp)], rtol=1e-6)


def test_merge(oracle):
    db = make_iid(3000, 32, 15)
    q = make_iid(5, 32, 16)
    D, I = oracle.search_flat(db, q, 20)
    parts = [oracle.search_flat(db[s:e], q, 20, id_offset=s) for s, e in [(0, 1000), (1000, 1700), (1700, 3000)]]
    Dm, Im = oracle.topk_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    assert np.array_equal(Im, I) and np.array_equal(Dm, D)
