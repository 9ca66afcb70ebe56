"""sessionsimilaritysearch_b200 — B200-native (sm_100a) retrieval hot path of SessionSimilaritySearch.

Host side mirrors the reference's Python call surface; all arithmetic runs in libsss_b200.so.
"""
from ._lib import (METRIC_IP, METRIC_L2, MODE_BF16, MODE_EXACT, MODE_FP32, NORM_FT, NORM_NONE, NORM_TORCH, NORM_UTIL,
                   REDUCE_MAX, REDUCE_NONE, REDUCE_SUM)
from .index import IndexBinaryFlat, IndexFlatIP, IndexFlatL2, build_index, normalize, pack_sign_bits
from .encoder import BinarizeHead, NodeAsinEmbedding, SessionEncoder, cosine_matrix, gather_rows, masked_mean_pool
from .metrics import get_ave_score, get_score, score_matrix
from .votes import ItemLists, get_prediction_by_knn, item_vote
from .featurize import (FeatureCache, FlatSessions, QueryVocab, featurize_batch, featurize_group, flatten,
                        flatten_prefixes)

__all__ = ["cosine_matrix", "masked_mean_pool", "get_score", "get_ave_score", "score_matrix", "NodeAsinEmbedding", "gather_rows", "FeatureCache", "FlatSessions", "QueryVocab", "featurize_batch", "featurize_group", "flatten", "flatten_prefixes", "SessionEncoder", "BinarizeHead", "ItemLists", "item_vote", "get_prediction_by_knn", "IndexFlatIP", "IndexFlatL2", "IndexBinaryFlat", "build_index", "normalize", "pack_sign_bits",
           "METRIC_IP", "METRIC_L2", "MODE_EXACT", "MODE_FP32", "MODE_BF16", "NORM_NONE", "NORM_UTIL", "NORM_FT",
           "NORM_TORCH", "REDUCE_NONE", "REDUCE_MAX", "REDUCE_SUM"]
