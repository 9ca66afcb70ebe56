// select.cu — the streaming top-k machinery shared by every scan kernel.
//
// A scan kernel (fp32 CUDA-core, bf16 tcgen05, Hamming) only FILTERS: a row is appended to its query's
// candidate list when score > thr[q].  Between scan waves `refine` reduces each list to the k best
// (optionally re-scoring new entries in fixed-order fp32 and collapsing rows to their session), and raises
// thr[q] to the k-th best score.  Waves run in row order, so a later row that merely ties the k-th score
// can never displace it (ties go to the smaller id): the strict compare keeps the result exact.
// Replaces the heap/reservoir inside faiss' IndexFlat*.search (test_amazon_filterd.py:578) [recalled].
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace sss {

// ---- expand: tensor-core hit records -> per-query candidate lists -----------------------------------
__global__ void __launch_bounds__(128) expand_records_kernel(const HitRecord* __restrict__ rec,
                                                             const uint32_t* __restrict__ rec_cnt, int rec_cap,
                                                             int64_t row_limit, SelectState st) {
  const int region = blockIdx.x;
  uint32_t n = rec_cnt[region];
  if (n > (uint32_t)rec_cap) {
    if (threadIdx.x == 0) *st.overflow = 1;
    n = rec_cap;
  }
  // 4 threads per record: each takes 8 of the 32 scores (two float4 loads)
  const int sub = threadIdx.x & 3;
  for (uint32_t e = threadIdx.x >> 2; e < n; e += blockDim.x >> 2) {
    const HitRecord* r = rec + (size_t)region * rec_cap + e;
    const uint32_t q = r->q;
    const uint32_t row_base = r->row_base + sub * 8;
    const float thr = st.thr[q];
    const float4* v4 = reinterpret_cast<const float4*>(r->v) + sub * 2;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float4 v = v4[c];
      float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int64_t row = (int64_t)row_base + c * 4 + t;
        if (vv[t] > thr && row < row_limit) {
          uint32_t slot = atomicAdd(&st.cnt[q], 1u);
          if (slot < (uint32_t)st.cap) st.cand[(size_t)q * st.cap + slot] = pack_cand(score_key(vv[t]), (uint32_t)row);
        }
      }
    }
  }
}

int launch_expand_records(const HitRecord* rec, const uint32_t* rec_cnt, int n_regions, int rec_cap, int64_t row_limit,
                          SelectState st, cudaStream_t stream) {
  expand_records_kernel<<<n_regions, 128, 0, stream>>>(rec, rec_cnt, rec_cap, row_limit, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- rescore: exact fixed-order fp32 score of every NEW candidate (EXACT mode) ------------------------
// One block per query, 8 warps.  A warp takes 32 new candidates at a time: their rows are read with fully
// coalesced 128-byte requests (one row segment per request, lane = column) into a warp-private shared tile,
// then lane l walks ITS row in k-ascending order with a single accumulator — the same rounding sequence as
// the fp32 scan and the oracle — and rewrites the candidate's key in place.
constexpr int RS_WARPS = 8;
constexpr int RS_KC = 32;  // columns staged per round
__global__ void __launch_bounds__(RS_WARPS * 32) rescore_kernel(RefineArgs a, SelectState st) {
  extern __shared__ float rs_smem[];
  float* qs = rs_smem;                                    // [d_round] the query
  const int d_round = (a.d + RS_KC - 1) / RS_KC * RS_KC;
  float* tiles = rs_smem + d_round;                       // [RS_WARPS][32][RS_KC + 1]
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = st.cnt[q];
  const int n = c > (uint32_t)st.cap ? st.cap : (int)c;
  const int nr = (int)st.nret[q];
  if (n <= nr) return;
  for (int j = threadIdx.x; j < d_round; j += blockDim.x) qs[j] = j < a.d ? a.q_f32[(size_t)q * a.d + j] : 0.0f;
  __syncthreads();
  uint64_t* g = st.cand + (size_t)q * st.cap;
  float* tile = tiles + (size_t)warp * 32 * (RS_KC + 1);
  for (int base = nr + warp * 32; base < n; base += RS_WARPS * 32) {
    const int i = base + lane;
    const uint64_t v = i < n ? g[i] : 0ull;
    const uint32_t my_row = i < n ? cand_id(v) : 0u;
    float acc = 0.0f;
    for (int k0 = 0; k0 < a.d; k0 += RS_KC) {
      const int col = k0 + lane;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const uint32_t row = __shfl_sync(0xffffffffu, my_row, r);
        tile[r * (RS_KC + 1) + lane] = col < a.d ? a.db_f32[(size_t)row * a.d + col] : 0.0f;
      }
      __syncwarp();
      const float* mine = tile + lane * (RS_KC + 1);
      if (a.metric == 0) {
#pragma unroll
        for (int kk = 0; kk < RS_KC; ++kk) acc = __fmaf_rn(qs[k0 + kk], mine[kk], acc);
      } else {
#pragma unroll
        for (int kk = 0; kk < RS_KC; ++kk) {
          // padded columns: q = x = 0 -> t = 0 -> acc unchanged
          const float t = __fsub_rn(qs[k0 + kk], mine[kk]);
          acc = __fmaf_rn(t, t, acc);
        }
      }
      __syncwarp();
    }
    if (i < n) g[i] = pack_cand(score_key(a.metric == 0 ? acc : -acc), my_row);
  }
}

int launch_rescore(const RefineArgs& a, SelectState st, cudaStream_t stream) {
  const int d_round = (a.d + RS_KC - 1) / RS_KC * RS_KC;
  size_t smem = sizeof(float) * ((size_t)d_round + (size_t)RS_WARPS * 32 * (RS_KC + 1));
  SSS_REQUIRE(smem <= 96 * 1024, "embedding width too large for rescore_kernel");
  static size_t smem_set = 0;
  if (smem > smem_set) {
    SSS_CUDA_OK(cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  rescore_kernel<<<(unsigned)a.nq, RS_WARPS * 32, smem, stream>>>(a, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- refine -----------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_desc(uint64_t* e, int P) {
  for (int k2 = 2; k2 <= P; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t a = e[i], b = e[ixj];
          bool desc = (i & k2) == 0;
          if (desc ? (a < b) : (a > b)) {
            e[i] = b;
            e[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) refine_kernel(RefineArgs a, SelectState st) {
  extern __shared__ uint64_t e[];
  __shared__ int s_m;
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int cap = st.cap;
  const uint32_t c = st.cnt[q];
  const int n = c > (uint32_t)cap ? cap : (int)c;
  if (c > (uint32_t)cap && tid == 0) *st.overflow = 1;
  const int nr = (int)st.nret[q];
  if (n == nr) return;  // nothing new since the last refine
  int P = 2;
  while (P < n) P <<= 1;
  uint64_t* g = st.cand + (size_t)q * cap;
  for (int i = tid; i < P; i += blockDim.x) {
    uint64_t v = i < n ? g[i] : 0ull;
    if (a.reduce_max && i >= nr && i < n) v = pack_cand(cand_key(v), (uint32_t)a.row_seg[cand_id(v)]);
    e[i] = v;
  }
  if (tid == 0) s_m = 0;
  __syncthreads();
  if (a.reduce_max) {
    // session-major order: (~id) in the high word, key in the low word; descending sort groups a
    // session's entries with its best score first.
    for (int i = tid; i < P; i += blockDim.x) {
      uint64_t v = e[i];
      e[i] = v ? ((v << 32) | (v >> 32)) : 0ull;
    }
    __syncthreads();
    bitonic_desc(e, P);
    uint32_t headmask = 0;  // P <= 8192, 256 threads -> at most 32 entries per thread
    int t = 0;
    for (int i = tid; i < P; i += blockDim.x, ++t) {
      uint64_t v = e[i];
      bool head = v != 0ull && (i == 0 || (uint32_t)(e[i - 1] >> 32) != (uint32_t)(v >> 32));
      headmask |= (head ? 1u : 0u) << t;
    }
    __syncthreads();
    t = 0;
    for (int i = tid; i < P; i += blockDim.x, ++t) {
      uint64_t v = e[i];
      e[i] = ((headmask >> t) & 1u) ? ((v << 32) | (v >> 32)) : 0ull;
    }
    __syncthreads();
  }
  bitonic_desc(e, P);
  for (int i = tid; i < P; i += blockDim.x)
    if (e[i] != 0ull && (i == P - 1 || e[i + 1] == 0ull)) s_m = i + 1;
  __syncthreads();
  const int m = s_m < a.k ? s_m : a.k;
  for (int i = tid; i < m; i += blockDim.x) g[i] = e[i];
  if (tid == 0) {
    st.cnt[q] = m;
    st.nret[q] = m;
    st.thr[q] = m == a.k ? key_score(cand_key(e[a.k - 1])) - st.margin[q] : -INFINITY;
  }
}

int launch_refine(const RefineArgs& a, SelectState st, cudaStream_t stream) {
  if (a.rescore && launch_rescore(a, st, stream)) return 1;
  size_t smem = (size_t)st.cap * sizeof(uint64_t);
  SSS_REQUIRE(st.cap <= 8192, "candidate capacity too large for refine_kernel");
  static bool attr_done = false;
  if (!attr_done) {
    SSS_CUDA_OK(cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
    attr_done = true;
  }
  refine_kernel<<<(unsigned)a.nq, 256, smem, stream>>>(a, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- emit -------------------------------------------------------------------------------------------
__global__ void emit_kernel(SelectState st, int64_t nq, int k, int metric, int64_t id_offset, float* __restrict__ D,
                            int64_t* __restrict__ I) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * k) return;
  int64_t q = t / k;
  int j = (int)(t % k);
  if ((uint32_t)j < st.nret[q]) {
    uint64_t c = st.cand[(size_t)q * st.cap + j];
    float s = key_score(cand_key(c));
    D[t] = metric == 0 ? s : -s;
    I[t] = (int64_t)cand_id(c) + id_offset;
  } else {
    D[t] = metric == 0 ? -INFINITY : INFINITY;
    I[t] = -1;
  }
}

int launch_emit(SelectState st, int64_t nq, int k, int metric, int64_t id_offset, float* D, int64_t* I,
                cudaStream_t stream) {
  int64_t total = nq * k;
  if (total <= 0) return 0;
  emit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(st, nq, k, metric, id_offset, D, I);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- k-way merge of per-shard candidates (after the NCCL all-gather) ---------------------
// One block per query; (key desc, id asc) bitonic sort of n_shards*k (key, id64) pairs in shared memory.
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ cD, const int64_t* __restrict__ cI,
                                                         int n_shards, int64_t nq, int k, int metric, int P,
                                                         float* __restrict__ D, int64_t* __restrict__ I) {
  extern __shared__ uint64_t sm[];
  uint64_t* keys = sm;                 // [P] key in the high word (0 = empty)
  int64_t* ids = (int64_t*)(sm + P);   // [P]
  const int64_t q = blockIdx.x;
  const int total = n_shards * k;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t kk = 0;
    int64_t id = INT64_MAX;
    if (i < total) {
      int s = i / k, j = i % k;
      int64_t gid = cI[((int64_t)s * nq + q) * k + j];
      if (gid >= 0) {
        float sc = cD[((int64_t)s * nq + q) * k + j];
        kk = (uint64_t)score_key(metric == 0 ? sc : -sc) + 1ull;  // +1: keep 0 for "empty"
        id = gid;
      }
    }
    keys[i] = kk;
    ids[i] = id;
  }
  __syncthreads();
  for (int k2 = 2; k2 <= P; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t ka = keys[i], kb = keys[ixj];
          int64_t ia = ids[i], ib = ids[ixj];
          bool a_before_b = ka > kb || (ka == kb && ia < ib);
          bool desc = (i & k2) == 0;
          bool sw = desc ? !a_before_b && !(ka == kb && ia == ib) : a_before_b;
          if (sw) {
            keys[i] = kb; keys[ixj] = ka;
            ids[i] = ib; ids[ixj] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    if (j < P && keys[j] != 0ull) {
      float s = key_score((uint32_t)(keys[j] - 1ull));
      D[q * k + j] = metric == 0 ? s : -s;
      I[q * k + j] = ids[j];
    } else {
      D[q * k + j] = metric == 0 ? -INFINITY : INFINITY;
      I[q * k + j] = -1;
    }
  }
}

int launch_topk_merge(const float* cD, const int64_t* cI, int n_shards, int64_t nq, int k, int metric, float* D,
                      int64_t* I, cudaStream_t stream) {
  if (nq <= 0 || k <= 0) return 0;
  int total = n_shards * k;
  int P = 2;
  while (P < total) P <<= 1;
  size_t smem = (size_t)P * 16;
  SSS_REQUIRE(smem <= 96 * 1024, "n_shards * k too large for topk_merge_kernel");
  SSS_CUDA_OK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  topk_merge_kernel<<<(unsigned)nq, 256, smem, stream>>>(cD, cI, n_shards, nq, k, metric, P, D, I);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
