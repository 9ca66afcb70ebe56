"""CPU tests of the evaluation-metric row: the oracle against the fixture the reference's own
get_score produced, and the native host seqratio (sss_seqratio_pairs runs on host threads, no GPU needed) against
the oracle's restatement of python-Levenshtein's algorithm."""
import os

import numpy as np
import pytest

from oracle import metrics_oracle as mo
from sessionsimilaritysearch_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_golden.npz")


def golden_data():
    z = np.load(GOLD)
    rng = np.random.default_rng(int(z["seed_test"]) + 7)
    test = [synth.split_session(s, rng) for s in synth.make_sessions(int(z["n_test"]), int(z["seed_test"]))]
    train = synth.make_sessions(int(z["n_train"]), int(z["seed_train"]))
    return z, test, train


@pytest.mark.parametrize("sim", ["all_jaccard", "cur_jaccard", "all_product_type_score"])
def test_oracle_matches_reference_generated_scores(sim):
    z, test, train = golden_data()
    gt = mo.score_matrix(z["I"], test, train, sim)
    assert np.array_equal(gt.view(np.uint32), z["gt_" + sim].view(np.uint32))
    assert np.float32(np.mean(gt)) == z["mean_" + sim]


def test_seqratio_restatement_known_answers():
    # identical lists, disjoint lists, one substitution of similar strings, the empty cases
    assert mo.seqratio(["a", "b"], ["a", "b"]) == 1.0
    assert mo.seqratio([], []) == 1.0
    assert mo.seqratio(["abc"], []) == 0.0
    assert mo.seqratio(["abc"], ["xyz"]) == 0.0            # substitution cost 2*6/6 = 2 = delete + insert
    # one string differs in one character of four: d = 2, q = 2 * 2 / 8 = 0.5 -> (4 - 0.5) / 4
    assert mo.seqratio(["k", "abcd"], ["k", "abcx"]) == pytest.approx((4 - 0.5) / 4)
    assert mo.edit_distance_x1("kitten", "sitting") == 5   # 2 substitutions (2 each) + 1 insertion


@pytest.mark.parametrize("sim", ["all_query_score", "all_product_title_score"])
def test_native_seqratio_equals_the_oracle(sim):
    from sessionsimilaritysearch_b200 import metrics
    z, test, train = golden_data()
    I = z["I"][:, :12].copy()
    I[0, 0] = -1  # padding id -> 0
    got = metrics.score_matrix(I, test, train, sim, n_threads=3)
    exp = mo.score_matrix(np.where(I < 0, 0, I), test, train, sim)
    exp[0, 0] = 0.0
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    assert np.any(got > 0) and np.any(got < 1)


def test_native_seqratio_unicode_and_empty_strings():
    from sessionsimilaritysearch_b200 import metrics
    mk = lambda titles: [(0, 'c', None, 'B', 'p', 'b', t, i) for i, t in enumerate(titles)]
    test = [(mk(["café \U0001F600", None]), mk(["", "x"])), (mk([None, None]), [])]
    train = [mk(["cafe \U0001F600", "", "x"]), mk([None]), mk(["", ""])]
    I = np.array([[0, 1, 2], [2, 1, 0]], np.int64)
    got = metrics.score_matrix(I, test, train, "all_product_title_score")
    exp = mo.score_matrix(I, test, train, "all_product_title_score")
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
