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

// ---- refine ------------------------------------------------------------------------------------------
// One block per query.  Input: the retained entries [0, nret) of the query's list (final keys, ids as
// returned) plus the NEW row-level candidates of the last scan wave, which arrive either
//   * as list entries [nret, cnt) appended with atomics by the fp32 / Hamming scans, or
//   * as tensor-core hit records in the query's private sub-regions (a.rec != nullptr): each record holds 32
//     raw scores, re-filtered here against the same threshold the scan used.
// Steps:
//   1. group by session with a shared-memory hash table (owner = session, best = max key);
//   2. EXACT mode only: a new row whose tensor-core score is more than 2*margin below its session's best
//      cannot hold the session's exact maximum (|exact - bf16| <= margin), so only the others survive and are
//      re-scored: rows are fetched with coalesced 16-byte loads into a warp-private tile, then lane l walks ITS
//      row in k-ascending order with one accumulator — the rounding sequence of the fp32 scan and the oracle;
//   3. per-session max of the final keys, compaction, bitonic sort, keep the best k, raise the threshold.
// Two instantiations run back to back every wave: a small one (most queries, several blocks per SM) and a
// large one that picks up the queries the small one had to skip (done[q] != wave).
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

template <int NE, int SLOT_BITS, int RSW, int KC, int THREADS, int RMAX, bool LAST>
struct RefineCfg {
  static constexpr int kNE = NE;               // max entries (retained + new) per query and wave
  static constexpr int kSlots = 1 << SLOT_BITS;
  static constexpr int kRsw = RSW;             // re-scoring warps
  static constexpr int kKc = KC;               // tile columns
  static constexpr int kThreads = THREADS;
  static constexpr int kRmax = RMAX;           // max hit records per query and wave
  static constexpr bool kLast = LAST;          // nobody behind us: too-large inputs are an overflow
  static constexpr int kMaxSub = 512;          // max record sub-regions per query (2 * grid_x)
  static constexpr size_t smem_bytes(int d_round, bool rescore) {
    return (size_t)NE * 8 + (size_t)kSlots * 8 + (size_t)NE * 2 + (size_t)NE * 2 + (size_t)RMAX * 4 +
           (size_t)(kMaxSub + 1) * 4 + sizeof(float) * ((size_t)RSW * 32 * (KC + 1) + (rescore ? d_round : 0));
  }
};

template <class C>
__global__ void __launch_bounds__(C::kThreads) refine_kernel(RefineArgs a, SelectState st) {
  extern __shared__ __align__(16) unsigned char rf_smem[];
  uint64_t* ent = reinterpret_cast<uint64_t*>(rf_smem);                      // [NE]
  uint32_t* owner = reinterpret_cast<uint32_t*>(ent + C::kNE);               // [slots] session + 1
  uint32_t* best = owner + C::kSlots;                                        // [slots] max key
  uint16_t* slot = reinterpret_cast<uint16_t*>(best + C::kSlots);            // [NE]
  uint16_t* list = slot + C::kNE;                                            // [NE] survivors
  uint32_t* recptr = reinterpret_cast<uint32_t*>(list + C::kNE);             // [RMAX] global record index
  uint32_t* subpre = recptr + C::kRmax;                                      // [kMaxSub + 1] prefix of record counts
  float* tiles = reinterpret_cast<float*>(subpre + C::kMaxSub + 1);          // [RSW][32][KC+1]
  float* qs = tiles + C::kRsw * 32 * (C::kKc + 1);                           // [d_round]
  __shared__ int s_n, s_nsurv, s_H, s_skip;

  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (st.done[q] == a.wave) return;  // the small instantiation already handled this query
  const int nr = (int)st.nret[q];
  uint64_t* g = st.cand + (size_t)q * st.cap;
  const float margin = st.margin[q];
  const int d_round = (a.d + C::kKc - 1) / C::kKc * C::kKc;
  int n;

  if (tid == 0) {
    s_nsurv = 0;
    s_H = 0;
    s_skip = 0;
    s_n = nr;
  }
  for (int h = tid; h < C::kSlots; h += C::kThreads) {
    owner[h] = 0u;
    best[h] = 0u;
  }
  if (a.rec != nullptr) {
    // ---- new candidates from hit records
    const int nsub = a.rec_nsub;
    const size_t sub0 = (size_t)q * nsub;
    if (tid == 0) subpre[0] = 0u;
    __syncthreads();
    // counts -> exclusive prefix (nsub <= 512: one pass by warp 0 over chunks of 32)
    if (warp == 0) {
      uint32_t run = 0;
      for (int s0 = 0; s0 < nsub; s0 += 32) {
        uint32_t cnt = 0;
        if (s0 + lane < nsub) {
          cnt = a.rec_cnt[sub0 + s0 + lane];
          if (cnt > (uint32_t)kRecSubCap) {
            *st.overflow = 1;
            cnt = kRecSubCap;
          }
        }
        uint32_t inc = cnt;
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (s0 + lane < nsub) subpre[s0 + lane + 1] = run + inc;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    __syncthreads();
    const int R = (int)subpre[nsub];
    if (R == 0) return;  // nothing new since the last refine
    if (R > C::kRmax) {
      if (!C::kLast) return;  // leave it to the large instantiation
      if (tid == 0) *st.overflow = 1;
    }
    const int Rc = R < C::kRmax ? R : C::kRmax;
    for (int s = tid; s < nsub; s += C::kThreads) {
      const uint32_t b0 = subpre[s], b1 = subpre[s + 1];
      for (uint32_t j = b0; j < b1 && j < (uint32_t)C::kRmax; ++j)
        recptr[j] = (uint32_t)((sub0 + s) * kRecSubCap + (j - b0));
    }
    for (int i = tid; i < nr; i += C::kThreads) ent[i] = g[i];
    __syncthreads();
    const float thr = st.thr[q];
    for (int t = tid; t < Rc * 32; t += C::kThreads) {
      const HitRecord* r = a.rec + recptr[t >> 5];
      const float v = r->v[t & 31];
      const int64_t row = (int64_t)r->row_base + (t & 31);
      if (v > thr && row < a.row_limit) {
        const int pos = atomicAdd(&s_n, 1);
        if (pos < C::kNE) ent[pos] = pack_cand(score_key(v), (uint32_t)row);
      }
    }
    __syncthreads();
    n = s_n;
    if (n > C::kNE) {
      if (!C::kLast) return;
      if (tid == 0) *st.overflow = 1;
      n = C::kNE;
    }
  } else {
    // ---- new candidates from the list
    const uint32_t c = st.cnt[q];
    n = c > (uint32_t)st.cap ? st.cap : (int)c;
    if (c > (uint32_t)st.cap && tid == 0) *st.overflow = 1;
    if (n == nr) return;  // nothing new since the last refine
    if (n > C::kNE) {
      if (!C::kLast) return;
      if (tid == 0) *st.overflow = 1;
      n = C::kNE;
    }
    for (int i = tid; i < n; i += C::kThreads) ent[i] = g[i];
  }
  if (a.rescore)
    for (int j = tid; j < d_round; j += C::kThreads) qs[j] = j < a.d ? a.q_f32[(size_t)q * a.d + j] : 0.0f;
  __syncthreads();

  // 1. group by session
  for (int i = tid; i < n; i += C::kThreads) {
    const uint64_t v = ent[i];
    const uint32_t id = cand_id(v);
    const uint32_t sess = (i < nr || !a.reduce_max) ? id : (uint32_t)a.row_seg[id];
    uint32_t h = (sess * 2654435761u) >> (32 - (C::kSlots == 4096 ? 12 : 11));
    for (int probe = 0; probe < C::kSlots; ++probe) {
      const uint32_t prev = atomicCAS(&owner[h], 0u, sess + 1u);
      if (prev == 0u || prev == sess + 1u) break;
      h = (h + 1u) & (C::kSlots - 1);
    }
    atomicMax(&best[h], cand_key(v));
    slot[i] = (uint16_t)h;
  }
  __syncthreads();

  if (a.rescore) {
    // 2a. survivors among the new rows
    for (int i = nr + tid; i < n; i += C::kThreads) {
      const float b = key_score(cand_key(ent[i]));
      const float lo = key_score(best[slot[i]]) - 2.0f * margin;
      if (b >= lo) list[atomicAdd(&s_nsurv, 1)] = (uint16_t)i;
    }
    __syncthreads();
    // final keys only from here on: retained entries keep theirs, survivors get exact ones
    for (int h = tid; h < C::kSlots; h += C::kThreads) best[h] = 0u;
    __syncthreads();
    for (int i = tid; i < nr; i += C::kThreads) atomicMax(&best[slot[i]], cand_key(ent[i]));
    // 2b. exact fixed-order re-scoring of the survivors
    const int nsurv = s_nsurv;
    if (warp < C::kRsw) {
      constexpr int LPR = C::kKc / 4;   // lanes per row segment (float4 each)
      constexpr int RPI = 32 / LPR;     // rows per load instruction
      float* tile = tiles + (size_t)warp * 32 * (C::kKc + 1);
      const int rsub = lane / LPR, c4 = lane % LPR;
      for (int base = warp * 32; base < nsurv; base += C::kRsw * 32) {
        const int li = base + lane;
        const int i = li < nsurv ? (int)list[li] : -1;
        const uint32_t my_row = i >= 0 ? cand_id(ent[i]) : 0u;
        float acc = 0.0f;
        for (int k0 = 0; k0 < a.d; k0 += C::kKc) {
          const int col = k0 + c4 * 4;
#pragma unroll
          for (int t = 0; t < 32 / RPI; ++t) {
            const int r = t * RPI + rsub;
            const uint32_t row = __shfl_sync(0xffffffffu, my_row, r);
            const float* src = a.db_f32 + (size_t)row * a.d + col;
            float4 v;
            if (col + 3 < a.d && (a.d & 3) == 0) {
              v = *reinterpret_cast<const float4*>(src);
            } else {
              v.x = col + 0 < a.d ? src[0] : 0.0f;
              v.y = col + 1 < a.d ? src[1] : 0.0f;
              v.z = col + 2 < a.d ? src[2] : 0.0f;
              v.w = col + 3 < a.d ? src[3] : 0.0f;
            }
            float* dst = tile + r * (C::kKc + 1) + c4 * 4;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
          }
          __syncwarp();
          const float* mine = tile + lane * (C::kKc + 1);
          if (a.metric == 0) {
#pragma unroll
            for (int kk = 0; kk < C::kKc; ++kk) acc = __fmaf_rn(qs[k0 + kk], mine[kk], acc);
          } else {
#pragma unroll
            for (int kk = 0; kk < C::kKc; ++kk) {
              const float t = __fsub_rn(qs[k0 + kk], mine[kk]);  // padded columns: 0 - 0
              acc = __fmaf_rn(t, t, acc);
            }
          }
          __syncwarp();
        }
        if (i >= 0) atomicMax(&best[slot[i]], score_key(a.metric == 0 ? acc : -acc));
      }
    }
    __syncthreads();
  }

  // 3. one entry per session, best k of them
  for (int h = tid; h < C::kSlots; h += C::kThreads) {
    const uint32_t o = owner[h];
    if (o != 0u && best[h] != 0u) ent[atomicAdd(&s_H, 1)] = pack_cand(best[h], o - 1u);
  }
  __syncthreads();
  const int H = s_H;
  int P = 2;
  while (P < H) P <<= 1;
  for (int i = H + tid; i < P; i += C::kThreads) ent[i] = 0ull;
  __syncthreads();
  bitonic_desc(ent, P);
  const int m = H < a.k ? H : a.k;
  for (int i = tid; i < m; i += C::kThreads) g[i] = ent[i];
  if (tid == 0) {
    st.cnt[q] = m;
    st.nret[q] = m;
    st.thr[q] = m == a.k ? key_score(cand_key(ent[a.k - 1])) - margin : -INFINITY;
    st.done[q] = a.wave;
  }
}

using RefineSmall = RefineCfg<2048, 11, 4, 16, 256, 512, false>;
using RefineLarge = RefineCfg<4096, 12, 6, 32, 512, 4096, true>;

template <class C>
static int launch_refine_cfg(const RefineArgs& a, SelectState st, cudaStream_t stream, size_t* smem_set) {
  const int d_round = (a.d + C::kKc - 1) / C::kKc * C::kKc;
  const size_t smem = C::smem_bytes(d_round, a.rescore != 0);
  SSS_REQUIRE(smem <= 200 * 1024, "embedding width too large for refine_kernel");
  if (smem > *smem_set) {
    SSS_CUDA_OK(cudaFuncSetAttribute(refine_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *smem_set = smem;
  }
  refine_kernel<C><<<(unsigned)a.nq, C::kThreads, smem, stream>>>(a, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_refine(const RefineArgs& a, SelectState st, cudaStream_t stream) {
  SSS_REQUIRE(st.cap <= RefineLarge::kNE, "candidate capacity too large for refine_kernel");
  SSS_REQUIRE(a.k <= RefineLarge::kNE / 2, "k too large for refine_kernel");
  SSS_REQUIRE(a.rec == nullptr || a.rec_nsub <= RefineLarge::kMaxSub, "too many record sub-regions per query");
  static size_t smem_small = 0, smem_large = 0;
  if (a.k <= RefineSmall::kNE / 4 && launch_refine_cfg<RefineSmall>(a, st, stream, &smem_small)) return 1;
  return launch_refine_cfg<RefineLarge>(a, st, stream, &smem_large);
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
