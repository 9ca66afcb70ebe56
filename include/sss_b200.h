/*
 * sss_b200.h — C ABI of libsss_b200.so, the B200-native (sm_100a) retrieval hot path of
 * SessionSimilaritySearch.
 *
 * The reference has no FFI: its "interface" for this path is a handful of Python calls into faiss
 * and PyG (all file:line citations are into the reference tree):
 *
 *   build_index(emb, metric)          test_amazon_filterd.py:207-223   -> sss_index_create + sss_index_add
 *   index.add(x)                      test_amazon_filterd.py:214,217,220; fine_tune_ours.py:843,849
 *   index.search(x, K) -> (D, I)      test_amazon_filterd.py:578; fine_tune_ours.py:876,882 -> sss_index_search
 *   normalize(vec)                    util_amazon_filtered.py:28-31; fine_tune_ours.py:38-40 -> sss_normalize
 *   faiss.IndexBinaryFlat             fine_tune_ours.py:839-843,871-876 -> sss_binary_*
 *   get_prediction_by_knn             test_amazon_filterd.py:59-78     -> sss_item_vote
 *   encoder(data) (GNN + pooling)     model/model.py:279-351, model/gnn.py:64-81,193-217 -> sss_encoder_*
 *   gnn(x_dict, edge_index_dict), pooling(node_emb, data), get_node=True       -> sss_encoder_forward_ex
 *   text embedder's masked mean       model/NodeEmbedding.py:112-125   -> sss_masked_mean
 *   F.normalize(a) @ F.normalize(b).T fine_tune_ours.py:133,480,494    -> sss_cosine_matrix
 *   get_score / get_ave_score         fine_tune_ours.py:42-97,883-897  -> sss_pair_scores, sss_seqratio_pairs
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on failure; sss_last_error() then returns a
 *     message for the calling thread (the Python facade re-raises it as RuntimeError, mirroring the
 *     reference's `raise RuntimeError(...)` convention, test_amazon_filterd.py:222).
 *   - plain pointers and sizes only.  `*_on_device` says whether a buffer is a CUDA device pointer
 *     (on the handle's device) or host memory; the library owns only its handle and workspaces.
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*; NULL = legacy
 *     default stream).  Host-buffer calls synchronise that stream before returning (their results are
 *     on the host); device-buffer search calls synchronise it only to read back an 8-byte status word
 *     (candidate-list overflow -> automatic rerun with a safe schedule; kernel watchdog).
 *   - a search is replayed from a CUDA graph captured on its first call with a given (nq, k, mode); add /
 *     set_segments drop the captured graphs.  Tuning knobs (SSS_WAVE_GROWTH, SSS_WAVE_FIRST, SSS_NO_BOOTSTRAP,
 *     SSS_NO_LAZY, SSS_NO_GRAPH, SSS_SCAN_VARIANT) are read from the environment ONCE, when a handle is created.
 *   - handles are not thread-safe; one handle per GPU for row-sharded search (several handles on several GPUs
 *     may live in one process).
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef SSS_B200_H
#define SSS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sss_index sss_index_t;
typedef struct sss_binary_index sss_binary_index_t;
typedef struct sss_encoder sss_encoder_t;

/* metric of a flat index (faiss.IndexFlatIP / IndexFlatL2, test_amazon_filterd.py:211-220) */
enum { SSS_METRIC_IP = 0, SSS_METRIC_L2 = 1 };

/* row normalisation applied by sss_index_add / sss_normalize */
enum {
  SSS_NORM_NONE = 0,
  SSS_NORM_UTIL = 1,  /* v / sqrt(clip(sum v^2, 1e-6))   util_amazon_filtered.py:28-31 */
  SSS_NORM_FT = 2,    /* v / (||v||_2 + 1e-4)            fine_tune_ours.py:38-40 */
  SSS_NORM_TORCH = 3  /* v / max(||v||_2, 1e-12)         F.normalize, fine_tune_ours.py:133,480,494 */
};

/* search mode */
enum {
  SSS_MODE_EXACT = 0, /* tcgen05 bf16 filter + fixed-order fp32 rescoring: ids and scores bit-identical to FP32 */
  SSS_MODE_FP32 = 1,  /* CUDA-core fixed-order fp32 FMA scan (k-ascending), the bit-faithful restatement */
  SSS_MODE_BF16 = 2   /* tcgen05 bf16 filter with a STATISTICAL slack (4 standard deviations of the bf16 rounding noise,
                         never more than EXACT's rigorous bound) + the same fp32 rescoring of what passes: returned
                         scores are fixed-order fp32, recall@k >= 0.999 against FP32 on embedding-like rows; up to
                         ~10x fewer rows re-scored than EXACT at d = 1600 */
};

/* per-session reduction of subsession (row) scores; segments are contiguous row ranges */
enum { SSS_REDUCE_NONE = 0, SSS_REDUCE_MAX = 1, SSS_REDUCE_SUM = 2 };

const char* sss_last_error(void);
/* library/ABI version, and the compute capability the kernels were built for (100 = sm_100a) */
int sss_version(void);
int sss_built_for_sm(void);

/* ---- flat index (replaces faiss.IndexFlatIP / IndexFlatL2) --------------------------------------- */

/* d: embedding width (any d >= 1).  id_offset: added to every returned id (row shard base). */
int sss_index_create(sss_index_t** out, int device, int d, int metric, int64_t id_offset);
int sss_index_destroy(sss_index_t* ix);

/* Append n rows of fp32 [n, d] (row-major).  norm_mode is applied on the GPU before storing.
 * Replaces index.add(normalize(emb)) / index.add(emb). */
int sss_index_add(sss_index_t* ix, const float* rows, int64_t n, int rows_on_device, int norm_mode,
                  void* stream);

/* Declare contiguous segments (sessions): seg_off[n_seg+1] (host, int64), seg_off[0]==0,
 * seg_off[n_seg]==ntotal.  reduce = MAX: score(session) = max over its rows; SUM: sum over its rows
 * (computed as <q, sum of rows>, rows summed in row order).  Returned ids are then session indices.
 * `stream` must be the stream the rows were added on (or one ordered after it); the call returns after it has
 * drained.  A rejected call leaves the index unchanged. */
int sss_index_set_segments(sss_index_t* ix, const int64_t* seg_off, int64_t n_seg, int reduce, void* stream);

int64_t sss_index_ntotal(const sss_index_t* ix);
int sss_index_dim(const sss_index_t* ix);

/* Top-k search.  q: fp32 [nq, d].  D: fp32 [nq, k] best first (IP: descending inner product;
 * L2: ascending squared distance).  I: int64 [nq, k].  Ties: smaller id first.  If fewer than k
 * results exist the tail is (IP: -inf, L2: +inf, id -1).  Replaces index.search(x, K). */
int sss_index_search(sss_index_t* ix, const float* q, int64_t nq, int k, int mode, int q_on_device,
                     float* D, int64_t* I, int out_on_device, void* stream);

/* Top-k search of one row shard for the sharded path: the same search, results written as ONE packed device block
 * [ids int64 nq*k | scores fp32 nq*k | pad | 16-byte trailer: int32 status] of sss_packed_bytes(nq, k) bytes — the
 * unit every rank contributes to the single all-gather; sss_topk_merge_packed consumes the gathered
 * blocks as they are.  async = 0: like sss_index_search (status read back, overflow re-run inside the call).
 * async = 1: nothing is read back and the call returns as soon as the work is enqueued; the search's status word
 * travels in the trailer, sss_topk_merge_packed ORs the trailers of all shards into *status_out, and a caller that
 * finds it non-zero (on every rank alike, since all ranks merge the same blocks) repeats the search with async = 0.
 * That removes the one host round trip in the middle of the sharded step. */
int64_t sss_packed_bytes(int64_t nq, int k);
int sss_index_search_packed(sss_index_t* ix, const float* q, int64_t nq, int k, int mode, int q_on_device,
                            void* packed, int async, void* stream);

/* Counters of the last search on this handle.  what: 0 = kernels launched, 1 = scan waves,
 * 2 = overflow reruns, 3 = scan-kernel device time in ns, 4 = scan-kernel launches, 5..8 = refine volumes
 * (candidates, rows re-scored, sessions sorted, refine invocations) — 3..8 need sss_index_set_profiling(ix, 1):
 * the search then runs as plain launches with CUDA events around every scan launch on the caller's stream instead of
 * replaying its captured graph —, 24 = bit mask of what overflowed when the last search had to be redone (1 record
 * sub-region, 2 records per query, 4 candidate list, 8 new candidates, 16 session table, 32 re-score list),
 * 25 = scan kernel of the last search (0 fp32 CUDA cores, 1 SS, 2 TS, 3 pair, 4 K-loop pair), 26 = 1 when the last
 * search replayed a captured CUDA graph.  Anything else: -1. */
int64_t sss_index_stat(const sss_index_t* ix, int what);
int sss_index_set_profiling(sss_index_t* ix, int on);

/* Row L2 normalisation of fp32 [n, d] into out (may alias in).  Replaces normalize(). */
int sss_normalize(const float* in, float* out, int64_t n, int d, int norm_mode, int on_device, int device,
                  void* stream);

/* k-way merge of per-shard candidates after the all-gather: cand_D fp32 [n_shards, nq, k],
 * cand_I int64 [n_shards, nq, k] (device) -> D [nq, k], I [nq, k] (device).  Same order rule as
 * sss_index_search; ids < 0 are padding. */
int sss_topk_merge(const float* cand_D, const int64_t* cand_I, int n_shards, int64_t nq, int k, int metric,
                   float* D, int64_t* I, int device, void* stream);
/* The same merge over n_shards packed blocks laid end to end (the output of all-gathering sss_index_search_packed
 * blocks).  Limit of both forms: n_shards * k <= 8192 (one shared-memory sort per query). */
int sss_topk_merge_packed(const void* gathered, int n_shards, int64_t nq, int k, int metric, float* D, int64_t* I,
                          int* status_out /* device int32, may be NULL */, int device, void* stream);

/* ---- binary index (replaces faiss.IndexBinaryFlat, fine_tune_ours.py:839-843,871-876) ------------ */

/* nbits must be a multiple of 8, <= 512; codes are uint8 [n, nbits/8] as produced by np.packbits(axis=1).
 * Codes of up to 256 bits (the reference hashes to 250) are searched on the tensor cores: the index also keeps
 * every code as +-1.0 in E4M3 (one byte per bit), <a, b> = nbits - 2 * hamming is exact in the fp32 accumulators,
 * and the fused scan + streaming top-k of the float index applies unchanged (bootstrap thresholds, waves, graph
 * replay).  Calls with at most 16 queries, and longer codes, take a popcount scan over the packed codes (an eighth of
 * the bytes of the tensor path's one-byte-per-bit rows). */
int sss_binary_create(sss_binary_index_t** out, int device, int nbits, int64_t id_offset);
int sss_binary_destroy(sss_binary_index_t* ix);
int sss_binary_add(sss_binary_index_t* ix, const uint8_t* codes, int64_t n, int on_device, void* stream);
int64_t sss_binary_ntotal(const sss_binary_index_t* ix);
/* the counters of sss_index_stat / the switch of sss_index_set_profiling, for a binary index */
int64_t sss_binary_stat(const sss_binary_index_t* ix, int what);
int sss_binary_set_profiling(sss_binary_index_t* ix, int on);
/* D: int32 [nq, k] Hamming distances ascending, I: int64 [nq, k]; ties: smaller id first. */
int sss_binary_search(sss_binary_index_t* ix, const uint8_t* q, int64_t nq, int k, int q_on_device,
                      int32_t* D, int64_t* I, int out_on_device, void* stream);

/* sign-binarise + pack: x fp32 [n, nbits_in] -> codes uint8 [n, ceil(nbits_in/8)], bit = (x > 0),
 * MSB first, zero padded: the (d+1)/2 -> astype(int) -> np.packbits of fine_tune_ours.py:839-840 applied
 * to BinarizeHead's {-1,0,+1} output (model/model.py:137). */
int sss_pack_sign_bits(const float* x, uint8_t* codes, int64_t n, int nbits_in, int on_device, int device,
                       void* stream);

/* ---- item vote (replaces get_prediction_by_knn, test_amazon_filterd.py:59-78) -------------------- */

/* For each of nq queries: neighbours I[q, 0..s) with weights D[q, 0..s); every item of neighbour
 * session j (items[item_off[j] .. item_off[j+1])) receives weight D[q, j]; weights are summed per item in arrival
 * order in float64 (the reference's defaultdict loop over a float64 array, :70-74); the top-K items by weight
 * descending, equal weights in order of first arrival (Python's stable sort, :76), are written to out_items
 * [nq, K] (-1 padded) and out_w [nq, K] (the float64 sums rounded to fp32).  All buffers on the device.
 * max_votes: an upper bound on the votes of one query (s * longest item list) or 0 if unknown — it only sizes the
 * shared-memory table.  Limits: K <= 256; item ids in [0, 2^32 - 2]; at most 12288 DISTINCT items voted for by one
 * query (the number of votes is unbounded; the reference's own call, sample_size = 500 x <= 19 items, needs 9500). */
int sss_item_vote(const float* D, const int64_t* I, int64_t nq, int s, const int64_t* item_off,
                  const int64_t* items, int64_t n_sessions, int K, int64_t max_votes, int64_t* out_items,
                  float* out_w, int device, void* stream);

/* ---- session encoder (replaces UnifyPoolingGraphLevelEncoder.forward after the text embedder) ----- */

/* Shapes of the reference model (pretrain_filtered_amazon.py:262-287): in_dim 768, hidden 800,
 * layers 3, out 1600, max_seq_len 20.  Weights are fp32, row-major [out, in] as in torch state_dicts.
 * in_dim and hidden must be multiples of 8 (the layers read 16-byte aligned slices of one bf16 operand buffer). */
typedef struct sss_encoder_shape {
  int in_dim;      /* 768 */
  int hidden;      /* 800 */
  int n_layers;    /* 3 */
  int out_dim;     /* 1600 */
  int max_seq_len; /* 20 */
} sss_encoder_shape_t;

int sss_encoder_create(sss_encoder_t** out, int device, const sss_encoder_shape_t* shape);
int sss_encoder_destroy(sss_encoder_t* enc);
/* Upload one named parameter (names follow the reference state_dict), fp32, host or device. */
int sss_encoder_set_param(sss_encoder_t* enc, const char* name, const float* data, int64_t numel, int on_device,
                          void* stream);

/* One batch of session graphs in CSR-free COO form, exactly the attributes the reference reads from a
 * PyG HeteroDataBatch (model/model.py:283-286,317; model/gnn.py:199-206).  All pointers are device. */
typedef struct sss_graph_batch {
  int64_t n_graphs;
  int64_t n_query;         /* query nodes */
  int64_t n_product;       /* distinct product nodes */
  int64_t n_expanded;      /* sum(cnt) product occurrences */
  const float* x_query;    /* [n_query, in_dim] text features */
  const float* x_product;  /* [n_product, in_dim] */
  const int64_t* query_batch;    /* [n_query] graph id */
  const int64_t* product_batch;  /* [n_product] */
  const int64_t* query_pos;      /* [n_query] pos_emb_id */
  const int64_t* product_cnt;    /* [n_product] */
  const int64_t* product_pos;    /* [n_expanded] pos_emb_id grouped by product */
  int64_t e_qp; const int64_t* qp_src; const int64_t* qp_dst; /* ('query','clicks','product') */
  int64_t e_pq; const int64_t* pq_src; const int64_t* pq_dst; /* ('product','clicked by','query') */
  int64_t e_pp; const int64_t* pp_src; const int64_t* pp_dst; /* ('product','to','product') */
} sss_graph_batch_t;

/* out: fp32 [n_graphs, out_dim] (device).  nonfinite (device int32, may be NULL) is set to 1 if any
 * input feature is NaN (the reference's isnan asserts, model/model.py:301-314, without host syncs). */
int sss_encoder_forward(sss_encoder_t* enc, const sss_graph_batch_t* batch, float* out, int32_t* nonfinite,
                        void* stream);

/* The two stages of the forward as the reference exposes them: gnn(x_dict, edge_index_dict) -> node embeddings
 * (model/gnn.py:64-81) and pooling(node_emb_dict, data) -> [B, out_dim] (model/gnn.py:193-217), and the node
 * embeddings of encoder(data, get_node=True) (model/model.py:344-351).
 *   run_gnn = 1: HeteroGGNN over batch->x_query / x_product; the node embeddings [n, in_dim + n_layers * hidden]
 *                (input features first, model/gnn.py:75-78) are written to z_query / z_product when non-NULL.
 *   run_gnn = 0: z_query / z_product are INPUTS of the pooling stage (x_query / x_product and the edges are unused).
 *   run_pooling = 1: PositionalAttentionPooling into out. */
typedef struct sss_encoder_io {
  float* out;
  float* z_query;
  float* z_product;
  int run_gnn;
  int run_pooling;
  int32_t* nonfinite; /* as in sss_encoder_forward; may be NULL */
} sss_encoder_io_t;
int sss_encoder_forward_ex(sss_encoder_t* enc, const sss_graph_batch_t* batch, const sss_encoder_io_t* io, void* stream);

/* Arithmetic of the encoder's dense linears: this library's own tcgen05 GEMM (csrc/gemm_bf16x3_sm100.cu) — operands
 * split into hi + lo bf16, three products accumulated in fp32 in TMEM; through the whole encoder 2.7e-5 of the output
 * scale from a float64 forward — with the work that follows each linear (attention logits, GRU gates, tanh / positional
 * concat, gated attention) fused into its epilogue.  It is the only arithmetic: sss_encoder_set_math accepts
 * SSS_ENCODER_MATH_BF16X3 and rejects the cuBLAS modes of earlier versions (there is no library GEMM behind this ABI
 * any more).  sss_encoder_stat(enc, 0) = kernels launched by the last forward. */
enum { SSS_ENCODER_MATH_FP32 = 0, SSS_ENCODER_MATH_BF16X9 = 1, SSS_ENCODER_MATH_BF16X3 = 2 };
int sss_encoder_set_math(sss_encoder_t* enc, int math);
int sss_encoder_get_math(const sss_encoder_t* enc);
int64_t sss_encoder_stat(const sss_encoder_t* enc, int what);

/* Row gather on the device: out[i, :] = table[ids[i], :] (fp32 [n_rows, d], int64 ids [n], out [n, d]).  Replaces the
 * nn.Embedding lookup of NodeAsinEmbedding.forward (model/NodeEmbedding.py:137-138) and serves the text-feature cache
 * of the batched featuriser.  An id outside [0, n_rows) fails with torch's own message ("index out of range in self"). */
int sss_gather_rows(const float* table, int64_t n_rows, int d, const int64_t* ids, int64_t n, float* out, int device,
                    void* stream);

/* ---- batched host featuriser (replaces the per-session Python of sequence_to_graph, util_amazon_filtered.py:98-230,
 * followed by PyG's Batch.from_data_list, test_amazon_filterd.py:485-488, for what the encoder reads) ---------- */

/* Sessions as flat HOST arrays.  An action is a search (act_is_search = 1, act_key = the caller's id of the query
 * string: the row of its text feature) or an item event (0, act_key = item id, the last field of the reference's
 * action tuple).  uniq_items lists every session's distinct item ids in node order — the reference takes
 * list(set(ids)), so the Python caller passes exactly that. */
typedef struct sss_flat_sessions {
  int64_t n_sessions;
  const int64_t* act_off;        /* [n_sessions + 1] */
  const uint8_t* act_is_search;  /* [n_actions] */
  const int64_t* act_key;        /* [n_actions] */
  const int64_t* uniq_off;       /* [n_sessions + 1] */
  const int64_t* uniq_items;     /* [sum of distinct items] */
} sss_flat_sessions_t;

/* Output: the batch-level arrays of sss_graph_batch_t (indices already batch-global), HOST buffers with the
 * capacities cap_* (sss_featurize_sizes gives the exact sizes).  pp_weight and last_click_mask may be NULL. */
typedef struct sss_graph_arrays {
  int64_t cap_query, cap_product, cap_expanded, cap_qp, cap_pp;
  int64_t n_query, n_product, n_expanded, e_qp, e_pp;  /* filled */
  int64_t* query_key;      /* [n_query] text key of the node (root_query_key for node 0 of every session) */
  int64_t* query_pos;      /* [n_query] pos_emb_id */
  int64_t* query_batch;    /* [n_query] session index */
  int64_t* product_key;    /* [n_product] item id (0 = the placeholder of an item-less session) */
  int64_t* product_cnt;    /* [n_product] */
  int64_t* product_batch;  /* [n_product] */
  int64_t* product_pos;    /* [n_expanded] pos_emb_id of every occurrence, grouped by product */
  int64_t* qp_src; int64_t* qp_dst;   /* [e_qp] ('query','clicks','product'); its transpose is 'clicked by' */
  int64_t* pp_src; int64_t* pp_dst;   /* [e_pp] ('product','to','product'), de-duplicated */
  float* pp_weight;                   /* [e_pp] multiplicity of the transition */
  float* last_click_mask;             /* [n_product] */
} sss_graph_arrays_t;

int sss_featurize_sizes(const sss_flat_sessions_t* s, int64_t* n_query, int64_t* n_product, int64_t* n_expanded,
                        int64_t* e_qp, int64_t* e_pp);
/* n_threads <= 0: all hardware threads (sessions are independent). */
int sss_featurize_batch(const sss_flat_sessions_t* s, int64_t root_query_key, sss_graph_arrays_t* out, int n_threads);
/* Many encoder batches in one call: the sessions form consecutive batches of batch_size sessions (the reference's
 * DataLoader(batch_size=200), test_amazon_filterd.py:488).  The output arrays hold the batches end to end, every node
 * and graph index LOCAL to its batch; bounds[(n_batches + 1) * 5] receives, per batch boundary, the offsets of
 * (query nodes, product nodes, expanded positions, q->p edges, p->p edges).  Sizes: sss_featurize_sizes. */
int sss_featurize_batches(const sss_flat_sessions_t* s, int64_t batch_size, int64_t root_query_key, sss_graph_arrays_t* out,
                          int64_t* bounds, int n_threads);

/* BinarizeHead eval forward, mlp=None (model/model.py:117-138): out = sign(x W^T + b) in {-1,0,+1}.
 * x [n, in], W [out, in], b [out], out [n, out]; device pointers. */
int sss_binarize_head(const float* x, const float* W, const float* b, int64_t n, int in_dim, int out_dim,
                      float* out, int device, void* stream);

/* ---- text-embedder tail (model/NodeEmbedding.py:112-125) ------------------------------------------------ */

/* Masked mean over tokens: out[n, :] = sum_t tok[n, t, :] * mask[n, t] / sum_t mask[n, t] — what
 * PretrainedQAEAEncoder.__call__ applies to the transformer's last_hidden_state.  tok fp32 [n, L, H], mask int64
 * [n, L] (the tokenizer's attention_mask), out fp32 [n, H]; device pointers.  A row whose mask is all zero gives NaN,
 * like the reference's 0 / 0. */
int sss_masked_mean(const float* tok, const int64_t* mask, int64_t n, int L, int H, float* out, int device, void* stream);

/* ---- in-batch cosine matrix (fine_tune_ours.py:133,480,494,613,626) ------------------------------------ */

/* out[i, j] = <a_i / max(||a_i||, 1e-12), b_j / max(||b_j||, 1e-12)>  (F.normalize(a) @ F.normalize(b).T);
 * a fp32 [na, d], b fp32 [nb, d], out fp32 [na, nb]; device pointers; fixed-order fp32 FMA. */
int sss_cosine_matrix(const float* a, int64_t na, const float* b, int64_t nb, int d, float* out, int device,
                      void* stream);

/* ---- evaluation metrics on the retrieved ids (get_score / get_ave_score, fine_tune_ours.py:42-97,883-897) ---- */

/* Ragged int64 lists per session in CSR form (device pointers): side a = the nq query sessions, side b = the database
 * sessions; I int64 [nq, k] = retrieved ids; out fp32 [nq, k] = the reference's `gt` matrix (ids outside [0, n_b): 0).
 * SSS_SCORE_JACCARD: lists hold each session's DISTINCT item ids (get_item: a set) -> |A & B| / |A | B|, 0 when both
 *   are empty ('all_jaccard' over prefix + suffix, 'cur_jaccard' over the prefix only: the caller picks the lists).
 * SSS_SCORE_TYPE_COSINE: lists hold the product-type ids of the item events in session order (get_item_type) ->
 *   cosine of the count vectors in float64 with numpy's summation order ('all_product_type_score'); at most 64
 *   distinct types per pair. */
enum { SSS_SCORE_JACCARD = 0, SSS_SCORE_TYPE_COSINE = 1 };
int sss_pair_scores(int kind, const int64_t* a_off, const int64_t* a_vals, int64_t nq, const int64_t* b_off,
                    const int64_t* b_vals, int64_t n_b, const int64_t* I, int k, float* out, int device, void* stream);

/* Lists of strings per session (HOST memory): strings seq_off[i] .. seq_off[i+1] belong to session i, string t is
 * the UTF-32 code points chars[str_off[t] .. str_off[t+1]). */
typedef struct sss_string_seqs {
  int64_t n_seqs;
  const int64_t* seq_off;
  const int64_t* str_off;
  const uint32_t* chars;
} sss_string_seqs_t;
/* out[q, j] = Levenshtein.seqratio(a[q], b[I[q, j]]) ('all_product_title_score'; with zero_if_empty = 1 the
 * 'all_query_score' rule: 0 when either list is empty).  Native host code on n_threads threads (<= 0: all):
 * string-sequence edit distance is nested dynamic programming over code points, not GPU work. */
int sss_seqratio_pairs(const sss_string_seqs_t* a, const sss_string_seqs_t* b, const int64_t* I, int64_t nq, int k,
                       int zero_if_empty, float* out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* SSS_B200_H */
