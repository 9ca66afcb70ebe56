// kernels.h — host-side launchers of every kernel in libsss_b200 (all return 0 / non-zero + set_error).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sss {

// prep.cu
int launch_add_rows(const float* in, int64_t n, int d, int d_pad, int norm_mode, float* out_f32, void* out_bf16,
                    int64_t out_row0, unsigned int* maxnorm2_bits, cudaStream_t st, int aug = 0);
// stats[0] = bits of max ||row||^2, stats[1] = bits of max ||row - bf16(row)||^2, stats[2] = bits of max ||row||_4^4
// (all from launch_add_rows)
// q_keep (optional, fp32 [nq_pad, d]): a private copy of the queries for the re-scoring passes (the caller's buffer
// is only read by this kernel, which is what lets a captured search graph re-bind one pointer per call)
// slack: 0 none, 1 rigorous bound (EXACT), 2 statistical bound (BF16) — see prep.cu
int launch_prep_queries(const float* q, int64_t nq, int64_t nq_pad, int d, int d_pad, void* q_bf16, int slack,
                        const unsigned int* stats, SelectState st, cudaStream_t stream, int aug = 0,
                        float* q_keep = nullptr);
int launch_gather_rows(const float* table, int64_t n_rows, int d, const int64_t* ids, int64_t n, float* out, int* bad,
                       cudaStream_t st);
int launch_row_seg(const int64_t* seg_off, int64_t n_seg, int32_t* row_seg, cudaStream_t st);
int launch_segment_sum(const float* rows, const int64_t* seg_off, int64_t n_seg, int d, float* out, cudaStream_t st);
int launch_normalize(const float* in, float* out, int64_t n, int d, int norm_mode, cudaStream_t st);

// scan_fp32.cu — CUDA-core fixed-order scan of rows [row_begin, row_end)
int launch_scan_fp32(const float* db, int d, int metric, int64_t row_begin, int64_t row_end, const float* q, int64_t nq,
                     SelectState st, cudaStream_t stream);

// scan_bf16_sm100.cu — TMA + tcgen05 scan of rows [row_begin, row_end) (row_begin multiple of 128)
struct Bf16ScanPlan {
  int num_kb;       // 64-wide K blocks (d_pad / 64)
  int num_mt;       // resident query m-tiles per CTA (1..4)
  int total_mtiles; // m-tiles of the whole padded query batch
  int num_stages;   // DB tile ring depth
  int grid_x;       // CTAs along the DB
  int grid_y;       // CTA groups along the queries
  int smem_bytes;
  int rec_cap;      // hit records per epilogue warp
  int n_regions;    // record sub-regions: nq_pad * grid_x * 2
  bool ts;          // query tiles resident in TMEM (A operand from TMEM) instead of shared memory
  bool two_cta;     // cta_group::2 pairs: M=256 x N=256 MMAs, each CTA loads half of every DB tile
  int rec_nsub;     // record sub-regions per query
  int tile_rows;    // DB rows per tile (128, 256 for the 2-CTA variant, 512 for the K-loop variant)
  bool kloop;       // wide rows (d_pad > 128): both operands streamed per K block, one query group per CTA pair
  int groups;       // K-loop variant: query groups of 256
  int last_k4;      // K-loop variant: 16-wide slices of real columns in the last K block
  bool fp8;         // operands are E4M3 bytes (Hamming search, +-1 codes): set by the caller after planning; rows of
                    // 2 * d_pad bytes either way, so tiles, tensor maps and the ring are those of the bf16 scan
};
// rec_boost multiplies the records per sub-region (1, or 4 after a search that overflowed one)
// d_used = columns that hold data (d, + 2 for L2); 0 = d_pad
// variant: 0 automatic (pair kernel above 128 queries, TS below), 1 SS, 2 TS, 3 pair — a tuning knob of the handle
int plan_scan_bf16(int d_pad, int64_t nq_pad, int num_sms, int max_stages, Bf16ScanPlan* plan, int rec_boost = 1,
                   int d_used = 0, int variant = 0);
// tensor maps are CUtensorMap objects (128 bytes each) built by make_tensor_map_2d
int make_tensor_map_bf16_2d(void* out_map128, const void* base, uint64_t rows, uint64_t cols_pad, uint32_t box_rows);
int launch_scan_bf16(const Bf16ScanPlan& plan, const void* tmap_q, const void* tmap_db, const void* q_bf16,
                     int64_t row_begin, int64_t row_end, SelectState st, HitRecord* rec, uint32_t* rec_cnt,
                     int* err_flag, float* cmax, cudaStream_t stream);
// gemm_bf16x3_sm100.cu — split-bf16 tensor-core GEMM of the encoder with fused epilogues: C[M,N] = A[M,K] * B[N,K]^T
// row_map (optional, device): output row r of the split takes source row row_map[r] (< 0: a zero row)
int launch_split_bf16(const float* x, int rows, int cols, int64_t ld, int transposed, void* hi, void* lo, int rows_pad,
                      int cols_pad, cudaStream_t stream, const int* row_map = nullptr);
enum { EPI_STORE = 0, EPI_ATT = 1, EPI_GRU = 2, EPI_POOL = 3, EPI_ATTPOOL = 4 };
struct GemmProblem {
  // operands: hi / lo bf16, K-major, row pitch in elements (a multiple of 64); A is read from column a_k0 on
  const void *a_hi, *a_lo;
  int a_rows_pad, a_ld, a_k0;
  const void *b_hi, *b_lo;
  int b_rows_pad, b_ld;
  int M, N;          // real output rows / columns (N only bounds the stores)
  int bn, tiles_n;   // N tile (128, or 96 for EPI_GRU) and their number
  int num_kb, last_k4;  // 64-wide K blocks; 16-wide slices of real columns in the last one (1..4)
  int epi;
  float* C;          // EPI_STORE / EPI_ATT: fp32 output; EPI_GRU: next features (fp32)
  int ldc;
  const float* bias;  // EPI_STORE (optional), EPI_POOL, EPI_ATTPOOL
  int sign_out;       // EPI_STORE: numerically sign(v) (BinarizeHead)
  // EPI_ATT: N is a sequence of parts of att_tiles_per_part tiles; part p < 4 with att[p] != NULL gets, per row and
  // tile, the partial <C[row, tile's columns], att[p]> in att_out[p][row * att_tiles_per_part + tile]
  const float* att[4];
  float* att_out[4];
  int att_tiles_per_part, att_width;
  // EPI_GRU
  const float* gru_gh; int gru_gh_ld;
  const float *gru_b_ih, *gru_b_hh;
  const float* gru_x; int gru_x_ld, gru_in_w;
  const float* gru_gp;
  int gru_H;
  // hi / lo bf16 copies of what the epilogue produces (EPI_GRU: columns out_k0 + unit; EPI_POOL: columns as in U)
  __nv_bfloat16 *out_hi, *out_lo;
  int out_ld, out_k0;
  // EPI_POOL
  int pool_is_product, pool_row0, pool_lin_w, pool_msl;
  const int* pool_prefix;     // products: occurrence rows [prefix[p], prefix[p + 1])
  const int64_t* pool_pos;    // pos_emb_id per occurrence (products) / per node (queries)
  const float* pool_pe;       // [msl, msl]
  float* pool_U;              // [n_expanded + n_query, lin_w + msl]
  // EPI_ATTPOOL
  const float* ap_bc;         // [n_graphs, N]
  const int* ap_node_graph;
  const float* ap_w;
  float* ap_out;              // [M, tiles_n] partial attention logits
};
// one launch = one or two independent problems (their tiles are enumerated back to back)
int launch_gemm_bf16x3(const GemmProblem* problems, int n_problems, int* err_flag, cudaStream_t stream);
// select.cu — bootstrap thresholds from chunk maxima: thr[q] = just below (2k-th largest chunk max - slack)
int launch_bootstrap_thr(const float* cmax, int n_chunks, int64_t nq, int64_t nq_pad, int k, int chunk_gap,
                         float slack_mult, SelectState st, cudaStream_t stream);

// select.cu
struct RefineArgs {
  int64_t nq;
  int k;
  int reduce_max;            // dedupe new entries by segment, keep the best
  const int32_t* row_seg;    // [n_rows] when reduce_max
  int rescore;               // recompute new entries' scores in fixed-order fp32
  const float* db_f32;       // [n_rows, d] when rescore
  const float* q_f32;        // [nq, d] when rescore
  int d;
  int metric;
  uint32_t wave;             // id of this wave (skip-list parity), > 0
  int all_large;             // set by launch_refine: k too large for the small instantiation
  const HitRecord* rec;      // tensor-core hit records of this wave, or nullptr for list input
  const uint32_t* rec_cnt;   // [nq_pad * rec_nsub]
  int rec_nsub;              // record sub-regions per query (2 * scan grid_x)
  int rec_cap;               // records per sub-region (16; 64 for the K-loop scan)
  int l2_tensor;             // L2 metric searched on the tensor path: thresholds live in tensor-score space
                             // ((||q||^2 - dist) / 2), keys in -dist space; refine converts at the boundary
  int lazy;                  // EXACT mode: keep candidate rows with tensor-core keys, re-score in the last wave
  int final;                 // last wave of the search
  int64_t row_limit;         // rows >= row_limit in a record are TMA zero fill
  unsigned long long* debug; // optional [12]: sums of candidates, re-scored rows, sessions, refines, then cycles per refine phase
};
// group by session, (EXACT) prune + re-score survivors, keep the best k, raise the threshold
int launch_refine(const RefineArgs& a, SelectState st, int num_sms, cudaStream_t stream);
int launch_emit(SelectState st, int64_t nq, int k, int metric, int64_t id_offset, float* D, int64_t* I, int* status_out,
                cudaStream_t stream);
// shard s: scores at cD + s * stride_d, ids at cI + s * stride_i (element strides)
// status_in (optional): one status word per shard, status_stride ints apart; their OR is written to status_out
int launch_topk_merge(const float* cD, const int64_t* cI, int64_t stride_d, int64_t stride_i, int n_shards, int64_t nq,
                      int k, int metric, float* D, int64_t* I, cudaStream_t stream, const int* status_in = nullptr,
                      int64_t status_stride = 0, int* status_out = nullptr);
// addresses of the two kernels whose pointer arguments a captured search graph re-binds per call (api.cu)
const void* emit_kernel_addr();
const void* prep_queries_kernel_addr();

// binary.cu
int launch_pack_sign_bits(const float* x, uint8_t* codes, int64_t n, int nbits, cudaStream_t st);
constexpr int kHammingSmallNq = 16;  // up to this many queries a binary search streams the PACKED codes (popcount scan)
int launch_scan_hamming(const uint8_t* db, int nbytes, int64_t row_begin, int64_t row_end, const uint8_t* q, int64_t nq,
                        SelectState st, cudaStream_t stream);
// dot_bits > 0: scores are +-1 dot products over dot_bits elements (tensor path); 0: scores are -hamming (popcount path)
int launch_emit_hamming(SelectState st, int64_t nq, int k, int dot_bits, int64_t id_offset, int32_t* D, int64_t* I,
                        cudaStream_t stream);
// packed codes [n, nbytes] -> +-1.0 E4M3 bytes [n, row_bytes] (row_bytes a multiple of 128, zero padded)
int launch_expand_codes_fp8(const uint8_t* codes, int nbytes, int64_t n, int row_bytes, uint8_t* out, cudaStream_t st);
// query staging of a binary search: fp8 rows [nq_pad, row_bytes] (may be NULL), packed rows [nq, pitch] (may be NULL),
// selection state reset
int launch_prep_binary(const uint8_t* q_codes, int64_t nq, int64_t nq_pad, int nbytes, int row_bytes, uint8_t* q_fp8,
                       uint8_t* q_packed, int pitch, SelectState st, cudaStream_t stream);
const void* prep_binary_kernel_addr();
const void* emit_hamming_kernel_addr();

}  // namespace sss
