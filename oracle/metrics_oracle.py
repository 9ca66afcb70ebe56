"""CPU restatement of the reference's evaluation metrics on the retrieved ids (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may use anything under oracle/).

    get_score(data_a, data_b, sim_type)                     fine_tune_ours.py:42-87
    get_ave_score(I, test_data, train_data, sim_type)       fine_tune_ours.py:89-96
    helpers get_item / get_item_type / get_query / get_session_item_title
                                                            util_amazon_filtered.py:33-37,59-60,234-238

Pinning: tests/golden/metrics_golden.npz holds scores produced by the reference's OWN get_score (source lines exec'd
by tests/golden/gen_metrics_golden.py) for the three sim types that are pure Python + numpy (all_jaccard,
cur_jaccard, all_product_type_score).  The two Levenshtein.seqratio types depend on python-Levenshtein (unpinned in
dependency.txt, not installable here): `seqratio` below restates its published algorithm (lev_edit_seq_distance over
lev_edit_distance with xcost = 1) from memory — **parity unpinned** for those two.
"""
import numpy as np


def get_item(session):
    return set([action[-1] for action in session if action[1] != 's'])


def get_session_item_title(session):
    return [action[-2] if action[-2] is not None else '' for action in session if action[1] != 's']


def get_item_type(session):
    return [action[4] for action in session if action[1] != 's' if action[4] is not None]


def get_query(sess, pad=True):
    q = [action[2] for action in sess if action[1] == 's' and action[2] is not None]
    return q if pad is False else [""] + q


def edit_distance_x1(a, b):
    """lev_edit_distance(..., xcost=1): insertions and deletions cost 1, a substitution 2 [recalled]"""
    if len(a) > len(b):
        a, b = b, a
    row = list(range(len(a) + 1))
    for j in range(1, len(b) + 1):
        prev, row[0] = row[0], j
        for i in range(1, len(a) + 1):
            cur = row[i]
            row[i] = min(prev + (0 if a[i - 1] == b[j - 1] else 2), row[i] + 1, row[i - 1] + 1)
            prev = cur
    return row[len(a)]


def edit_seq_distance(s1, s2):
    """lev_edit_seq_distance [recalled]: edit distance of two string SEQUENCES; replacing string x by y costs
    2 * d(x, y) / (len x + len y) (<= 1 + 1), insert / delete cost 1.  Includes the library's quirk: when both
    strings of a cell are empty the inner pointer is not advanced."""
    s1, s2 = list(s1), list(s2)
    while s1 and s2 and s1[0] == s2[0]:
        s1.pop(0)
        s2.pop(0)
    while s1 and s2 and s1[-1] == s2[-1]:
        s1.pop()
        s2.pop()
    if not s1:
        return float(len(s2))
    if not s2:
        return float(len(s1))
    if len(s1) > len(s2):
        s1, s2 = s2, s1
    n1, n2 = len(s1) + 1, len(s2) + 1
    row = [float(i) for i in range(n2)]
    for i in range(1, n1):
        x1 = s1[i - 1]
        D = i - 1.0
        x = float(i)
        j2 = 0  # the library's len2p / str2p
        for p in range(1, n2):
            y = s2[j2]
            l = len(x1) + len(y)
            if l == 0:
                q = D
            else:
                d = edit_distance_x1(x1, y)
                j2 += 1
                q = D + 2.0 / l * d
            x += 1.0
            if x > q:
                x = q
            D = row[p]
            if x > D + 1.0:
                x = D + 1.0
            row[p] = x
    return row[n2 - 1]


def seqratio(a, b):
    lensum = len(a) + len(b)
    if lensum == 0:
        return 1.0
    return (lensum - edit_seq_distance(a, b)) / lensum


def get_score(data_a, data_b, sim_type):
    if sim_type == 'all_jaccard':
        a_item = get_item(data_a[0] + data_a[1])
        b_item = get_item(data_b[0] + data_b[1])
        return len(a_item & b_item) / len(a_item | b_item)
    if sim_type == 'cur_jaccard':
        a_item = get_item(data_a[0])
        b_item = get_item(data_b[0])
        c = len(a_item | b_item)
        return 0 if c == 0 else len(a_item & b_item) / c
    if sim_type == 'all_query_score':
        a_query = get_query(data_a[0] + data_a[1], pad=False)
        b_query = get_query(data_b[0] + data_b[1], pad=False)
        if len(a_query) == 0 or len(b_query) == 0:
            return 0
        return seqratio(a_query, b_query)
    if sim_type == 'all_product_title_score':
        return seqratio(get_session_item_title(data_a[0] + data_a[1]), get_session_item_title(data_b[0] + data_b[1]))
    if sim_type == 'all_product_type_score':
        a_type = get_item_type(data_a[0] + data_a[1])
        b_type = get_item_type(data_b[0] + data_b[1])
        vec_len = len(set(a_type + b_type))
        type_to_id = {}
        a_vec = np.zeros(vec_len)
        b_vec = np.zeros(vec_len)
        for t in a_type:
            if t not in type_to_id:
                type_to_id[t] = len(type_to_id)
            a_vec[type_to_id[t]] += 1
        if len(a_type) > 0:
            a_vec = a_vec / np.linalg.norm(a_vec)
        for t in b_type:
            if t not in type_to_id:
                type_to_id[t] = len(type_to_id)
            b_vec[type_to_id[t]] += 1
        if len(b_type) > 0:
            b_vec = b_vec / np.linalg.norm(b_vec)
        return np.sum(a_vec * b_vec)
    raise RuntimeError("unrecognized sim type: %s" % sim_type)


def score_matrix(I, test_data, train_data, sim_type):
    """gt of fine_tune_ours.py:883-888 / get_ave_score: float32 [nq, K]"""
    gt = np.zeros_like(I, dtype=np.float32)
    for i, t in enumerate(test_data):
        for j, d in enumerate(I[i, :]):
            gt[i, j] = get_score(t, (train_data[d], []), sim_type)
    return gt


def get_ave_score(I, test_data, train_data, sim_type):
    return np.mean(score_matrix(I, test_data, train_data, sim_type))
