"""Host featuriser: session action tuples -> SessionGraph, with the signature and the output layout of the
reference's sequence_to_graph (util_amazon_filtered.py:98-230 and its helpers :7-22, :33-37, :62-95).

An action is (ts, type, keyword, asin, ptype, brand, title, item_id); type == 's' is a search, anything else an
item event (field order from decompose_data.py:13,17).

Layout produced:
  query nodes    node 0 is the empty-string root, then one node per search in order;
                 pos_emb_id = len(seq) - [0, i+1 for the search at position i]
  product nodes  distinct item ids in list(set(...)) order; cnt = occurrences; pos_emb_id = len(seq) - j for every
                 occurrence j, grouped by product; an item-less session gets one node (id 0, cnt 1, pos 0, 'UNK')
  edges          ('query','clicks','product'): one per item event from the latest search node (multi-edges kept),
                 plus its transpose; ('product','to','product'): consecutive item transitions, de-duplicated with a
                 count weight, self transitions kept
  also           last_click_mask, query_target / product_target / text token tensors, 'ori_seq'
"""
import torch

from .graph import EDGE_PP, EDGE_PQ, EDGE_QP, SessionGraph

SEARCH = 's'


def _is_search(action):
    return action[1] == SEARCH


def _tok(tokenizer, strings, max_length):
    t = tokenizer(strings, padding='max_length', max_length=max_length, truncation=True, return_tensors="pt")
    return t['input_ids'], t['token_type_ids'], t['attention_mask']


def get_query_node_tokens(session_details, tokenizer, max_length):
    words, where = [""], [0]
    for i, act in enumerate(session_details):
        if _is_search(act):
            words.append(act[2] if act[2] is not None else "")
            where.append(i + 1)
    ids, tti, am = _tok(tokenizer, words, max_length)
    return ids, tti, am, len(session_details) - torch.tensor(where)


def get_item(session):
    return set(act[-1] for act in session if not _is_search(act))


def get_all_query(seq):
    return [act[2] for act in seq if _is_search(act) and act[2] is not None]


def get_next_query(seq):
    q = get_all_query(seq)
    return q[0] if q else None


def get_query(sess, pad=True):
    q = get_all_query(sess)
    return ([""] + q) if pad else q


def get_session_item_title(session):
    return [act[-2] if act[-2] is not None else '' for act in session if not _is_search(act)]


def get_item_type(session):
    return [act[4] for act in session if not _is_search(act) and act[4] is not None]


def session_to_text(session):
    out = []
    for act in session:
        s = act[2] if _is_search(act) else act[-2]
        out.append("" if s is None else s)
    return out


def _occurrences(seq):
    """item id -> positions j of its events, in session order"""
    occ = {}
    for j, act in enumerate(seq):
        if not _is_search(act):
            occ.setdefault(act[-1], []).append(j)
    return occ


def get_item_title(seq, item_list):
    occ = _occurrences(seq)
    titles = []
    for item in item_list:
        if item in occ:
            t = seq[occ[item][0]][-2]
            titles.append("" if t is None else t)
    return titles


def get_item_pos_cnt(seq, item_list):
    occ = _occurrences(seq)
    pos, cnt = [], []
    for item in item_list:
        js = occ.get(item, [])
        cnt.append(len(js))
        pos.extend(len(seq) - j for j in js)
    return pos, cnt


def sequence_to_graph(idx, seq, tar, tokenizer, query_max_len, ignore_query=False):
    g = SessionGraph()
    g['idx'].idx = idx
    if ignore_query:
        seq = [act for act in seq if not _is_search(act)]

    # ---- query nodes
    q = g['query']
    q.x, q.token_type_ids, q.attention_mask, q.pos_emb_id = get_query_node_tokens(seq, tokenizer, query_max_len)
    q.input_ids = q.x
    q.num_nodes = q.x.shape[0]
    q.mask = torch.ones(q.num_nodes)
    q.mask[0] = 0

    # ---- future queries (training target)
    qt = g['query_target']
    future = get_all_query(tar)
    qt.mask = torch.ones(len(future)) if future else torch.zeros(1)
    future = future or [""]
    qt.input_ids, qt.token_type_ids, qt.attention_mask = _tok(tokenizer, future, query_max_len)
    qt.num_nodes = len(future)

    # ---- product nodes
    items = list(get_item(seq))
    pos_ids, counts = get_item_pos_cnt(seq, items)
    assert sum(counts) == len(pos_ids) and len(counts) == len(items)
    if not items:
        items, counts, pos_ids = [0], [1], [0]   # the "unknown product" placeholder
    slot = {it: i for i, it in enumerate(items)}
    p = g['product']
    p.x = torch.LongTensor(items)
    p.num_nodes = len(items)
    p.cnt = torch.tensor(counts)
    p.pos_emb_id = torch.tensor(pos_ids)
    titles = get_item_title(seq, items)
    if not titles:
        assert items == [0]
        titles = ['UNK']
    p.input_ids, p.token_type_ids, p.attention_mask = _tok(tokenizer, titles, query_max_len)
    assert p.input_ids.size(0) == p.num_nodes
    p.mask = torch.ones(p.num_nodes)
    g['ori_seq'] = (seq, tar)

    # ---- future products (training target)
    pt = g['product_target']
    tar_items = list(get_item(tar))
    pt.y = torch.LongTensor(tar_items)
    pt.num_nodes = n_t = len(tar_items)
    t_titles = get_item_title(tar, tar_items or [0]) or ['UNK']
    ids, tti, am = _tok(tokenizer, t_titles, query_max_len)
    pt.input_ids, pt.token_type_ids, pt.attention_mask = ids[:n_t], tti[:n_t], am[:n_t]
    pt.mask = torch.ones(n_t)
    assert pt.input_ids.size(0) == pt.num_nodes

    # ---- query -> product edges: every item event hangs off the latest search node
    cur, src, dst = 0, [], []
    for act in seq:
        if _is_search(act):
            cur += 1
            continue
        if act[3] is None and act[-1] != 0:
            raise RuntimeError("asin is None")
        src.append(cur)
        dst.append(slot[act[-1]])
    g[EDGE_QP].edge_index = torch.tensor([src, dst], dtype=torch.long)
    g[EDGE_PQ].edge_index = torch.tensor([dst, src], dtype=torch.long)
    g[EDGE_QP].edge_weight = None
    g[EDGE_PQ].edge_weight = None

    # ---- product -> product transitions, de-duplicated with multiplicity
    chain = [slot[act[-1]] for act in seq if not _is_search(act)] or [slot[0] if 0 in slot else 0]
    first_seen = {}
    last = 0
    for a, b in zip(chain[:-1], chain[1:]):
        first_seen[(a, b)] = first_seen.get((a, b), 0) + 1   # dicts keep first-insertion order
        last = b
    p.last_click_mask = torch.zeros_like(p.x).float()
    p.last_click_mask[last] = 1
    pairs = list(first_seen)
    g[EDGE_PP].edge_index = torch.tensor([[a for a, _ in pairs], [b for _, b in pairs]], dtype=torch.long)
    g[EDGE_PP].edge_weight = torch.tensor([first_seen[k] for k in pairs], dtype=torch.float32)
    p.y = p.x

    # ---- whole-session text
    tx = g['text']
    tx.input_ids, tx.token_type_ids, tx.attention_mask = _tok(tokenizer, [''] + session_to_text(seq), 20)
    tx.num_nodes = tx.input_ids.shape[0]
    return g
