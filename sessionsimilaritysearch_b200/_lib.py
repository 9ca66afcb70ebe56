"""ctypes binding of libsss_b200.so (include/sss_b200.h).  There is no fallback: if the CUDA library is
missing or was not built, importing a symbol raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsss_b200.so")

METRIC_IP, METRIC_L2 = 0, 1
NORM_NONE, NORM_UTIL, NORM_FT, NORM_TORCH = 0, 1, 2, 3
MODE_EXACT, MODE_FP32, MODE_BF16 = 0, 1, 2
REDUCE_NONE, REDUCE_MAX, REDUCE_SUM = 0, 1, 2
MODES = {"exact": MODE_EXACT, "fp32": MODE_FP32, "bf16": MODE_BF16}
REDUCES = {None: REDUCE_NONE, "none": REDUCE_NONE, "max": REDUCE_MAX, "sum": REDUCE_SUM}

c_i64, c_int, c_vp, c_fp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p


class EncoderShape(ctypes.Structure):
    _fields_ = [("in_dim", c_int), ("hidden", c_int), ("n_layers", c_int), ("out_dim", c_int), ("max_seq_len", c_int)]


class GraphBatch(ctypes.Structure):
    _fields_ = [("n_graphs", c_i64), ("n_query", c_i64), ("n_product", c_i64), ("n_expanded", c_i64),
                ("x_query", c_vp), ("x_product", c_vp), ("query_batch", c_vp), ("product_batch", c_vp),
                ("query_pos", c_vp), ("product_cnt", c_vp), ("product_pos", c_vp),
                ("e_qp", c_i64), ("qp_src", c_vp), ("qp_dst", c_vp),
                ("e_pq", c_i64), ("pq_src", c_vp), ("pq_dst", c_vp),
                ("e_pp", c_i64), ("pp_src", c_vp), ("pp_dst", c_vp)]


class EncoderIO(ctypes.Structure):
    _fields_ = [("out", c_vp), ("z_query", c_vp), ("z_product", c_vp), ("run_gnn", c_int), ("run_pooling", c_int),
                ("nonfinite", c_vp)]


class StringSeqs(ctypes.Structure):
    _fields_ = [("n_seqs", c_i64), ("seq_off", c_vp), ("str_off", c_vp), ("chars", c_vp)]


class FlatSessions(ctypes.Structure):
    _fields_ = [("n_sessions", c_i64), ("act_off", c_vp), ("act_is_search", c_vp), ("act_key", c_vp),
                ("uniq_off", c_vp), ("uniq_items", c_vp)]


class GraphArrays(ctypes.Structure):
    _fields_ = [("cap_query", c_i64), ("cap_product", c_i64), ("cap_expanded", c_i64), ("cap_qp", c_i64),
                ("cap_pp", c_i64), ("n_query", c_i64), ("n_product", c_i64), ("n_expanded", c_i64), ("e_qp", c_i64),
                ("e_pp", c_i64), ("query_key", c_vp), ("query_pos", c_vp), ("query_batch", c_vp),
                ("product_key", c_vp), ("product_cnt", c_vp), ("product_batch", c_vp), ("product_pos", c_vp),
                ("qp_src", c_vp), ("qp_dst", c_vp), ("pp_src", c_vp), ("pp_dst", c_vp), ("pp_weight", c_vp),
                ("last_click_mask", c_vp)]


# name -> (restype, argtypes); every symbol declared in include/sss_b200.h
SIGNATURES = {
    "sss_last_error": (ctypes.c_char_p, []),
    "sss_version": (c_int, []),
    "sss_built_for_sm": (c_int, []),
    "sss_index_create": (c_int, [ctypes.POINTER(c_vp), c_int, c_int, c_int, c_i64]),
    "sss_index_destroy": (c_int, [c_vp]),
    "sss_index_add": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp]),
    "sss_index_set_segments": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp]),
    "sss_index_ntotal": (c_i64, [c_vp]),
    "sss_index_dim": (c_int, [c_vp]),
    "sss_index_search": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "sss_packed_bytes": (c_i64, [c_i64, c_int]),
    "sss_index_search_packed": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "sss_topk_merge_packed": (c_int, [c_vp, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp]),
    "sss_index_stat": (c_i64, [c_vp, c_int]),
    "sss_index_set_profiling": (c_int, [c_vp, c_int]),
    "sss_normalize": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp]),
    "sss_topk_merge": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "sss_binary_create": (c_int, [ctypes.POINTER(c_vp), c_int, c_int, c_i64]),
    "sss_binary_destroy": (c_int, [c_vp]),
    "sss_binary_add": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp]),
    "sss_binary_ntotal": (c_i64, [c_vp]),
    "sss_binary_stat": (c_i64, [c_vp, c_int]),
    "sss_binary_set_profiling": (c_int, [c_vp, c_int]),
    "sss_binary_search": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "sss_pack_sign_bits": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp]),
    "sss_item_vote": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_int, c_vp]),
    "sss_encoder_create": (c_int, [ctypes.POINTER(c_vp), c_int, ctypes.POINTER(EncoderShape)]),
    "sss_encoder_destroy": (c_int, [c_vp]),
    "sss_encoder_set_param": (c_int, [c_vp, ctypes.c_char_p, c_vp, c_i64, c_int, c_vp]),
    "sss_encoder_forward": (c_int, [c_vp, ctypes.POINTER(GraphBatch), c_vp, c_vp, c_vp]),
    "sss_encoder_forward_ex": (c_int, [c_vp, ctypes.POINTER(GraphBatch), ctypes.POINTER(EncoderIO), c_vp]),
    "sss_masked_mean": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_int, c_vp]),
    "sss_cosine_matrix": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_int, c_vp]),
    "sss_pair_scores": (c_int, [c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_int, c_vp, c_int, c_vp]),
    "sss_seqratio_pairs": (c_int, [ctypes.POINTER(StringSeqs), ctypes.POINTER(StringSeqs), c_vp, c_i64, c_int, c_int,
                                   c_vp, c_int]),
    "sss_gather_rows": (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_int, c_vp]),
    "sss_featurize_sizes": (c_int, [ctypes.POINTER(FlatSessions), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                    ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "sss_featurize_batch": (c_int, [ctypes.POINTER(FlatSessions), c_i64, ctypes.POINTER(GraphArrays), c_int]),
    "sss_featurize_batches": (c_int, [ctypes.POINTER(FlatSessions), c_i64, c_i64, ctypes.POINTER(GraphArrays),
                                      ctypes.POINTER(c_i64), c_int]),
    "sss_encoder_set_math": (c_int, [c_vp, c_int]),
    "sss_encoder_get_math": (c_int, [c_vp]),
    "sss_encoder_stat": (c_i64, [c_vp, c_int]),
    "sss_binarize_head": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_int, c_vp]),
}

_lib = None


def load():
    """dlopen the in-tree library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libsss_b200.so is not built (%s). Run `python -m sessionsimilaritysearch_b200.build`; there is no "
            "CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().sss_last_error()
        raise RuntimeError(msg.decode("utf-8", "replace") if msg else "libsss_b200 call failed")


def current_device():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except ImportError:
        pass
    return 0


def current_stream(device=None):
    """cudaStream_t of torch's current stream on `device` (0 = legacy default stream without torch)."""
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_stream(device).cuda_stream)
    except ImportError:
        pass
    return 0
