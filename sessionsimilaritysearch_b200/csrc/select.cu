// select.cu — the streaming top-k machinery shared by every scan kernel.
//
// A scan kernel (fp32 CUDA-core, bf16 tcgen05, Hamming) only FILTERS: a row is appended to its query's
// candidate list when score > thr[q].  Between scan waves `refine` reduces each list to the k best
// (optionally re-scoring new entries in fixed-order fp32 and collapsing rows to their session), and raises
// thr[q] to the k-th best score.  Waves run in row order, so a later row that merely ties the k-th score
// can never displace it (ties go to the smaller id): the strict compare keeps the result exact.
// Replaces the heap/reservoir inside faiss' IndexFlat*.search (test_amazon_filterd.py:578) [recalled].
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace sss {

// ---- refine ------------------------------------------------------------------------------------------
// Per query and wave.  Input: the retained entries [0, nret) of the query's list plus the NEW row-level candidates of
// the last scan wave, which arrive either
//   * as list entries [nret, cnt) appended with atomics by the fp32 / Hamming scans, or
//   * as tensor-core hit records in the query's private sub-regions (a.rec != nullptr): each record holds 32
//     raw scores, re-filtered here against the threshold the scan used.
// Phases (all in shared memory):
//   A  compaction: sub-region counters (one load per thread) -> prefix; ONE THREAD PER RECORD loads its 144 bytes,
//      builds the pass mask and appends the passing rows with one shared-memory atomic per warp;
//   B  group by session in a hash table (owner = session, best = max key);
//   S  lazy representation only (EXACT mode, bootstrapped searches, k <= 256): the retained entries are ROWS with
//      tensor-core keys; a 3-pass radix select gives a lower bound of b_k (k-th best session) and the floor
//      b_k - 2 * margin below which a session can never decide the result;
//   C  survivors: rows of live sessions within 2 * margin of their session's best tensor-core score
//      (|exact - tensor| <= margin, so the others cannot hold the session's exact maximum);
//   W  lazy, not the last wave: write the survivors back as rows, thr = b_k - 2 * margin — done;
//   D  otherwise: exact fixed-order re-scoring of the survivors — one lane walks one fp32 row in k-ascending order
//      with a single accumulator, the rounding sequence of the fp32 scan and the oracle;
//   E  per-session max of the final keys, gather, sort, keep the best k (eager representation: one exact entry per
//      session), thr = k-th exact score - margin.
// A small instantiation (one block per query, 256 threads, ~38 KB, 4 blocks per SM) serves the common case; queries it
// cannot hold go on a skip list that a large instantiation (persistent, one block per SM) drains right after.
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

// Descending sort of n distinct non-zero keys in e[0, n) (zero padded to P by the caller for the bitonic path).
// Small inputs are sorted by rank counting: every thread reads all keys (shared-memory broadcasts), counts the
// larger ones and scatters its own key to that rank through `tmp` — two barriers instead of ~40.
__device__ __forceinline__ void sort_desc(uint64_t* e, uint64_t* tmp, int n, int P) {
  if (n <= 32 && tmp != nullptr) {  // rank counting only pays for tiny inputs (instruction count, measured)
    uint64_t mine[4];
    int rank[4];
    int cnt = 0;
    for (int i = threadIdx.x; i < n && cnt < 4; i += blockDim.x, ++cnt) {
      mine[cnt] = e[i];
      rank[cnt] = 0;
    }
    for (int j = 0; j < n; ++j) {
      const uint64_t v = e[j];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < cnt) rank[c] += v > mine[c] ? 1 : 0;
    }
    for (int c = 0; c < cnt; ++c) tmp[rank[c]] = mine[c];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) e[i] = tmp[i];
    __syncthreads();
    return;
  }
  bitonic_desc(e, P);
}

template <int SLOT_BITS, int NC, int SURV, int KMAX, int RSW, int KC, int THREADS, int RMAX, bool LAST>
struct RefineCfg {
  static constexpr int kSlotBits = SLOT_BITS;
  static constexpr int kSlots = 1 << SLOT_BITS;  // hash slots = max distinct sessions per query and wave
  static constexpr int kNc = NC;                 // max new candidates per query and wave
  static constexpr int kSurv = SURV;             // max rows re-scored per query and wave
  static constexpr int kKmax = KMAX;             // max k (retained entries)
  static constexpr int kRsw = RSW;               // (unused since the per-lane row walk)
  static constexpr int kKc = KC;                 // query staging is zero-padded to a multiple of this
  static constexpr int kThreads = THREADS;
  static constexpr int kRmax = RMAX;             // max hit records per query and wave
  static constexpr bool kLast = LAST;            // nobody behind us: too-large inputs are an overflow
  static constexpr int kTmpBytes = 1280;         // radix-select histogram (258 ints) / rank-sort staging (32 keys)
  static constexpr int kMaxSub = 512;            // max record sub-regions per query (2 * grid_x)
  // scratch = [recptr | subpre], reused as the gather/sort array A[kSlots] once the passes are done
  __host__ __device__ static constexpr size_t scratch_bytes() {
    size_t s = (size_t)RMAX * 4 + (size_t)(kMaxSub + 1) * 4 + 8;
    size_t a = (size_t)kSlots * 8;
    return ((s > a ? s : a) + 15) / 16 * 16;
  }
  // rows at least this wide are re-scored through a per-warp [32][kWideCols + 1] staging tile filled with cp.async
  // (coalesced 128-byte row segments, no registers held while the loads are in flight)
  static constexpr int kWideRow = 512;
  static constexpr int kWideCols = THREADS <= 256 ? 64 : 32;   // columns per step (the 16-warp instantiation has less room)
  static constexpr size_t tile_bytes(int d, bool rescore) {
    return rescore && d >= kWideRow ? (size_t)(THREADS / 32) * 32 * (kWideCols + 1) * sizeof(float) : 0;
  }
  static constexpr size_t smem_bytes(int d_round, bool rescore, int d = 0) {
    return scratch_bytes() + kTmpBytes + (size_t)kSlots * 8 + (size_t)NC * 8 + 16 + (size_t)NC * 2 + (size_t)SURV * 2 +
           (size_t)KMAX * 2 + 8 + 16 + sizeof(float) * (rescore ? d_round + 64 : 0) + tile_bytes(d, rescore);
  }
};

enum { RF_DONE = 0, RF_SKIP = 1 };

template <class C>
struct RefineSmem {
  uint64_t* A;          // [slots] gather + sort (overlays the scratch below)
  uint64_t* tmp;        // [512] rank-sort staging
  uint32_t* recptr;     // [RMAX] global record index
  uint32_t* subpre;     // [kMaxSub + 1]
  uint32_t* owner;      // [slots] session + 1
  uint32_t* best;       // [slots] max key
  uint32_t* ent_key;    // [NC] compacted new candidates
  uint32_t* ent_row;    // [NC]
  int* ctr;             // [4] n_ent, n_surv, H, uniq / fail
  uint16_t* ent_slot;   // [NC]
  uint16_t* surv;       // [SURV] entry index of the rows to re-score
  uint16_t* ret_slot;   // [KMAX]
  float* qs;            // [d_round (+ 64: zero padding up to a whole step of the staged walk)]
  float* tile;          // [warps][32][kWideCols + 1] staging of the wide-row re-scoring (present when d >= kWideRow)
  __device__ explicit RefineSmem(unsigned char* p, int d_round = 0) {
    A = reinterpret_cast<uint64_t*>(p);
    recptr = reinterpret_cast<uint32_t*>(p);
    subpre = recptr + C::kRmax;
    tmp = reinterpret_cast<uint64_t*>(p + C::scratch_bytes());
    unsigned char* rest = p + C::scratch_bytes() + C::kTmpBytes;
    owner = reinterpret_cast<uint32_t*>(rest);
    best = owner + C::kSlots;
    ent_key = best + C::kSlots;
    ent_row = ent_key + C::kNc;
    ctr = reinterpret_cast<int*>(ent_row + C::kNc);
    ent_slot = reinterpret_cast<uint16_t*>(ctr + 4);
    surv = ent_slot + C::kNc;
    ret_slot = surv + C::kSurv;
    qs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ret_slot + C::kKmax) + 15) & ~(uintptr_t)15);  // float4 reads
    tile = qs + d_round + 64;
  }
};

// claim-or-find the slot of `sess`; -1 when the table is full
template <class C>
__device__ __forceinline__ int rf_insert(const RefineSmem<C>& sm, uint32_t sess) {
  uint32_t h = (sess * 2654435761u) >> (32 - C::kSlotBits);
  for (int probe = 0; probe < C::kSlots; ++probe) {
    const uint32_t prev = atomicCAS(&sm.owner[h], 0u, sess + 1u);
    if (prev == 0u) {
      atomicAdd(&sm.ctr[3], 1);
      return (int)h;
    }
    if (prev == sess + 1u) return (int)h;
    h = (h + 1u) & (C::kSlots - 1);
  }
  atomicOr(&sm.ctr[3], 0x40000000);
  return -1;
}

// warp-aggregated append: lanes with `pass` get consecutive positions after *counter
__device__ __forceinline__ int warp_append(int* counter, bool pass) {
  const uint32_t bal = __ballot_sync(0xffffffffu, pass);
  if (bal == 0u) return -1;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == (__ffs(bal) - 1)) base = atomicAdd(counter, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
  return pass ? base + __popc(bal & ((1u << lane) - 1u)) : -1;
}

// per-phase cycle accounting of thread 0 (a.debug[4 + phase]): experiment builds only (-DSSS_EXPERIMENT); release
// kernels carry no clock reads
#ifdef SSS_EXPERIMENT
#define RF_PHASE_BEGIN() long long _t0 = clock64()
#define RF_PHASE(i)                                                              \
  do {                                                                           \
    if (a.debug != nullptr && threadIdx.x == 0) {                                \
      const long long _t = clock64();                                            \
      atomicAdd(&a.debug[4 + (i)], (unsigned long long)(_t - _t0));              \
      _t0 = _t;                                                                  \
    }                                                                            \
  } while (0)
#else
#define RF_PHASE_BEGIN() do { } while (0)
#define RF_PHASE(i) do { } while (0)
#endif

template <class C>
__device__ int refine_query(const RefineArgs& a, const SelectState& st, const RefineSmem<C>& sm, int q) {
  RF_PHASE_BEGIN();
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NW = C::kThreads / 32;
  // Lazy representation (EXACT mode, tensor-core waves only): the retained entries [0, nl) are candidate ROWS with
  // their tensor-core keys (bit 31 of nret marks it) and the exact re-scoring is deferred to the last wave.  With m
  // the filter slack (|exact - tensor| <= m) and b_k the k-th best session by tensor-core score: k sessions have an
  // exact score >= b_k - m, so a session whose best tensor-core score is below b_k - 2m, and inside a live session a
  // row more than 2m below the session's best, can never decide the result.  What survives that test is kept as
  // rows; the last wave (or a query whose band outgrows the list) re-scores the survivors once and continues in
  // the eager representation (one exact entry per session).
  const uint32_t nret_raw = st.nret[q];
  const int nr_all = (int)(nret_raw & 0x7FFFFFFFu);
  const bool lazy = a.lazy != 0 && a.rescore != 0 && a.rec != nullptr && ((nret_raw >> 31) != 0u || nr_all == 0);
  const int nr = lazy ? 0 : nr_all;   // retained exact entries
  const int nl = lazy ? nr_all : 0;   // retained lazy rows
  uint64_t* g = st.cand + (size_t)q * st.cap;
  // L2 on the tensor path: the scan (and st.thr, st.margin) live in tensor-score space, score = (||q||^2 - dist) / 2;
  // keys, the retained lists and every comparison below live in -dist space: key value = 2 * score - ||q||^2, slack
  // doubled.  Thresholds are converted back (rounded down) when they are published.
  const bool l2t = a.l2_tensor != 0;
  const float qn2 = l2t ? st.qn2[q] : 0.0f;
  const float thr_scan = st.thr[q];
  const float margin = l2t ? 2.0f * st.margin[q] : st.margin[q];
  const float thr = l2t ? __fmaf_rn(2.0f, thr_scan, -qn2) : thr_scan;
  auto to_scan = [&](float t) { return l2t ? __fmul_rd(0.5f, __fadd_rd(t, qn2)) : t; };
  const int d_round = (a.d + C::kKc - 1) / C::kKc * C::kKc;

  __syncthreads();  // previous query of a persistent block is fully done with shared memory
  for (int h = tid; h < C::kSlots; h += C::kThreads) {
    sm.owner[h] = 0u;
    sm.best[h] = 0u;
  }
  if (tid < 4) sm.ctr[tid] = tid == 0 ? nl : 0;
  if (tid == 0) sm.subpre[0] = 0u;
  if (nl > C::kKmax && !C::kLast) return RF_SKIP;
  for (int i = tid; i < nl; i += C::kThreads) {  // lazy rows re-enter as candidates in front of the new ones
    const uint64_t e = g[i];
    sm.ent_key[i] = cand_key(e);
    sm.ent_row[i] = cand_id(e);
  }
  __syncthreads();

  // ---- A. compact the new candidates into shared memory
  if (a.rec != nullptr) {
    const int nsub = a.rec_nsub;
    const size_t sub0 = (size_t)q * nsub;
    // record counts of the query's sub-regions: one load per thread, then an exclusive prefix by warp 0
    for (int s = tid; s < nsub; s += C::kThreads) {
      uint32_t cnt = a.rec_cnt[sub0 + s];
      if (cnt > (uint32_t)a.rec_cap) {
        atomicOr(st.overflow, 1);   // reason codes: sss_index_stat(ix, 24)
        cnt = (uint32_t)a.rec_cap;
      }
      sm.subpre[s + 1] = cnt;
    }
    __syncthreads();
    if (warp == 0) {
      uint32_t run = 0;
      for (int s0 = 0; s0 < nsub; s0 += 32) {
        uint32_t inc = s0 + lane < nsub ? sm.subpre[s0 + lane + 1] : 0u;
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (s0 + lane < nsub) sm.subpre[s0 + lane + 1] = run + inc;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    __syncthreads();
    RF_PHASE(0);  // init + sub-region counts
    const int R = (int)sm.subpre[nsub];
    if (R == 0 && !(lazy && a.final != 0 && nl > 0)) return RF_DONE;  // nothing new since the last refine
    if (R > C::kRmax) {
      if (!C::kLast) return RF_SKIP;
      if (tid == 0) atomicOr(st.overflow, 2);
    }
    const int Rc = R < C::kRmax ? R : C::kRmax;
    for (int s = tid; s < nsub; s += C::kThreads) {
      const uint32_t b0 = sm.subpre[s], b1 = sm.subpre[s + 1];
      for (uint32_t j = b0; j < b1 && j < (uint32_t)C::kRmax; ++j)
        sm.recptr[j] = (uint32_t)((sub0 + s) * (size_t)a.rec_cap + (j - b0));
    }
    __syncthreads();
    // one thread per record: its nine 16-byte loads are independent, so a block sees ONE global latency for all
    // of the query's records; the few scores that pass are re-read (L1) when they are appended
    for (int r0 = 0; r0 < Rc; r0 += C::kThreads) {
      const int ri = r0 + tid;
      uint32_t mask = 0u, row_base = 0u;
      const HitRecord* r = nullptr;
      if (ri < Rc) {
        r = a.rec + sm.recptr[ri];
        const uint4* rv = reinterpret_cast<const uint4*>(r->v);
        row_base = r->row_base;
        uint4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = rv[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          mask |= (__uint_as_float(x[i].x) > thr_scan ? 1u : 0u) << (4 * i);
          mask |= (__uint_as_float(x[i].y) > thr_scan ? 1u : 0u) << (4 * i + 1);
          mask |= (__uint_as_float(x[i].z) > thr_scan ? 1u : 0u) << (4 * i + 2);
          mask |= (__uint_as_float(x[i].w) > thr_scan ? 1u : 0u) << (4 * i + 3);
        }
        const int64_t room = a.row_limit - (int64_t)row_base;  // rows past the end of the index (last tile)
        if (room < 32) mask = room <= 0 ? 0u : (mask & ((1u << (int)room) - 1u));
      }
      const int mine = __popc(mask);
      int inc = mine;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      int base = 0;
      if (lane == 31 && inc > 0) base = atomicAdd(&sm.ctr[0], inc);
      base = __shfl_sync(0xffffffffu, base, 31);
      int pos = base + inc - mine;
      while (mask != 0u) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1u;
        if (pos < C::kNc) {
          const float v = r->v[bit];
          sm.ent_key[pos] = score_key(l2t ? __fmaf_rn(2.0f, v, -qn2) : v);
          sm.ent_row[pos] = row_base + (uint32_t)bit;
        }
        ++pos;
      }
    }
  } else {
    const uint32_t c = st.cnt[q];
    const int n = c > (uint32_t)st.cap ? st.cap : (int)c;
    if (c > (uint32_t)st.cap && tid == 0) atomicOr(st.overflow, 4);
    if (n == nr) return RF_DONE;  // nothing new since the last refine
    if (n - nr > C::kNc) {
      if (!C::kLast) return RF_SKIP;
    }
    for (int i = nr + tid; i < n; i += C::kThreads) {
      const uint64_t e = g[i];
      if (i - nr < C::kNc) {
        sm.ent_key[i - nr] = cand_key(e);
        sm.ent_row[i - nr] = cand_id(e);
      }
    }
    if (tid == 0) sm.ctr[0] = n - nr;
  }
  // retained entries own their slots and keep their (final) keys
  for (int i = tid; i < nr; i += C::kThreads) {
    const uint64_t e = g[i];
    const int h = rf_insert<C>(sm, cand_id(e));
    if (h >= 0) atomicMax(&sm.best[h], cand_key(e));
    sm.ret_slot[i] = (uint16_t)(h >= 0 ? h : 0);
  }
  __syncthreads();
  RF_PHASE(1);  // record compaction + retained insert
  int n_ent = sm.ctr[0];
  if (n_ent > C::kNc) {
    if (!C::kLast) return RF_SKIP;
    if (tid == 0) atomicOr(st.overflow, 8);
    n_ent = C::kNc;
  }

  // ---- B. group by session (four session gathers in flight per thread)
  for (int i0 = tid; i0 < n_ent; i0 += 4 * C::kThreads) {
    uint32_t sess[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * C::kThreads;
      const uint32_t row = sm.ent_row[i < n_ent ? i : i0];
      sess[u] = a.reduce_max ? (uint32_t)a.row_seg[row] : row;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * C::kThreads;
      if (i < n_ent) {
        const int h = rf_insert<C>(sm, sess[u]);
        if (h >= 0) atomicMax(&sm.best[h], sm.ent_key[i]);
        sm.ent_slot[i] = (uint16_t)(h >= 0 ? h : 0);
      }
    }
  }
  __syncthreads();
  RF_PHASE(2);  // session hash
  {
    const int uniq = sm.ctr[3];
    if ((uniq & 0x40000000) != 0 || (uniq & 0x3FFFFFFF) > C::kSlots * 3 / 4) {
      if (!C::kLast) return RF_SKIP;
      if ((uniq & 0x40000000) != 0 && tid == 0) atomicOr(st.overflow, 16);  // table full: candidates were dropped
    }
  }

  float t_lo = -INFINITY;  // session-level floor of the survivors (eager: none — it costs more than it saves)
  float new_thr = thr;
  if (lazy) {
    // ---- S. k-th best session by tensor-core score -> floor b_k - 2 * margin (rounded down: conservative).
    // Three 8-bit radix passes over the session table; the low 8 key bits are left zero, i.e. the result is a lower
    // bound of b_k within 2^-15 relative — any lower bound keeps the pruning valid.
    if ((sm.ctr[3] & 0x3FFFFFFF) >= a.k) {
      int* hist = reinterpret_cast<int*>(sm.tmp);  // [256] bins + [2] result
      uint32_t prefix = 0u;
      int rem = a.k;
      for (int pass = 0; pass < 3; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 256; i += C::kThreads) hist[i] = 0;
        __syncthreads();
        for (int h = tid; h < C::kSlots; h += C::kThreads) {
          const uint32_t key = sm.best[h];
          if (sm.owner[h] != 0u && (pass == 0 || (key >> (shift + 8)) == prefix))
            atomicAdd(&hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (warp == 0) {  // lane l owns bins 255-8l .. 248-8l (descending)
          int c[8], s = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            c[j] = hist[255 - 8 * lane - j];
            s += c[j];
          }
          int incl = s;
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          int run = incl - s;
          if (run < rem && rem <= incl) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (run < rem && rem <= run + c[j]) {
                hist[256] = 255 - 8 * lane - j;
                hist[257] = rem - run;
              }
              run += c[j];
            }
          }
        }
        __syncthreads();
        prefix = (prefix << 8) | (uint32_t)hist[256];
        rem = hist[257];
        __syncthreads();
      }
      t_lo = __fsub_rd(key_score(prefix << 8), 2.0f * margin);
      new_thr = fmaxf(thr, t_lo);
    }
    RF_PHASE(3);  // lazy: k-th best session
  }

  if (a.rescore) {
    // ---- C. survivors: rows of live sessions within 2 * margin of their session's best tensor-core score
    for (int i0 = warp * 32; i0 < n_ent; i0 += C::kThreads) {
      const int i = i0 + lane;
      bool keep = false;
      if (i < n_ent) {
        const float sb = key_score(sm.best[sm.ent_slot[i]]);
        keep = sb >= t_lo && key_score(sm.ent_key[i]) >= sb - 2.0f * margin;
      }
      const int pos = warp_append(&sm.ctr[1], keep);
      if (pos >= 0 && pos < C::kSurv) sm.surv[pos] = (uint16_t)i;
    }
    __syncthreads();
    int nsurv = sm.ctr[1];
    if (nsurv > C::kSurv) {
      if (!C::kLast) return RF_SKIP;
      if (tid == 0) atomicOr(st.overflow, 32);
      nsurv = C::kSurv;
    }
    if (lazy && a.final == 0) {
      if (nsurv <= C::kKmax) {
        // ---- W. stay lazy: the surviving rows with their tensor-core keys are the retained set
        for (int i = tid; i < nsurv; i += C::kThreads) {
          const int ei = (int)sm.surv[i];
          g[i] = pack_cand(sm.ent_key[ei], sm.ent_row[ei]);
        }
        if (tid == 0) {
          st.cnt[q] = (uint32_t)nsurv;
          st.nret[q] = nsurv > 0 ? ((uint32_t)nsurv | 0x80000000u) : 0u;
          st.thr[q] = fmaxf(thr_scan, to_scan(new_thr));  // (never lower it: the conversion rounds down)
          if (a.debug != nullptr) {
            atomicAdd(&a.debug[0], (unsigned long long)n_ent);
            atomicAdd(&a.debug[2], (unsigned long long)(sm.ctr[3] & 0x3FFFFFFF));
            atomicAdd(&a.debug[3], 1ull);
          }
        }
        RF_PHASE(4);
        return RF_DONE;
      }
      if (!C::kLast) return RF_SKIP;  // the band outgrew the small list: the large instantiation decides
    }
    for (int j = tid; j < d_round + 64; j += C::kThreads) sm.qs[j] = j < a.d ? a.q_f32[(size_t)q * a.d + j] : 0.0f;
    // final keys only from here on: retained entries keep theirs, survivors get exact ones
    for (int h = tid; h < C::kSlots; h += C::kThreads) sm.best[h] = 0u;
    __syncthreads();
    for (int i = tid; i < nr; i += C::kThreads) atomicMax(&sm.best[sm.ret_slot[i]], cand_key(g[i]));
    RF_PHASE(4);  // survivors + reset
    // ---- D. exact fixed-order re-scoring of the survivors: one lane walks one row in k-ascending order with a
    // single accumulator (the oracle's rounding sequence).  A lane streams its own row as 16-byte loads, 64 bytes
    // (two sectors) per step with the next step already in flight; the second half of every sector is an L1 hit
    // (bypassing L1 with ld.global.cg was measured 1.5x slower on 1600-wide rows: profiles/r01_ab_experiments.md).
    if (a.d >= C::kWideRow) {
      // wide rows (the 1600-wide session embeddings): a lane-per-row walk with 16-byte loads touches 32 different
      // rows per instruction and half a sector each time, and its loads are serialised by the registers they need.
      // Here a warp takes 32 survivors and fetches their rows kWideCols columns at a time with cp.async straight
      // into a [32][kWideCols + 1] shared-memory tile (lane = column: full 128-byte segments, 32-64 copies in flight
      // per lane, no registers held); every lane then runs ITS row's products in k-ascending order into its single
      // accumulator — the same rounding sequence as the scalar walk and the oracle.
      constexpr int W = C::kWideCols, P = W + 1;
      float* tile = sm.tile + warp * (32 * P);
      const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
      for (int base = warp * 32; base < nsurv; base += C::kThreads) {
        const int li = base + lane;
        const bool have = li < nsurv;
        const int ei = (int)sm.surv[have ? li : base];
        const uint32_t my_row = sm.ent_row[ei];
        float acc = 0.0f;
        for (int c0 = 0; c0 < a.d; c0 += W) {
          __syncwarp();  // the previous step's readers are done
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const uint32_t row_r = __shfl_sync(0xffffffffu, my_row, r);
            const float* src = a.db_f32 + (size_t)row_r * a.d + c0 + lane;
#pragma unroll
            for (int cc = 0; cc < W / 32; ++cc) {
              const uint32_t dst = tile_s + (uint32_t)((r * P + cc * 32 + lane) * 4);
              if (c0 + cc * 32 + lane < a.d)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + cc * 32) : "memory");
              else
                tile[r * P + cc * 32 + lane] = 0.0f;
            }
          }
          asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
          __syncwarp();
#pragma unroll 8
          for (int j2 = 0; j2 < W; ++j2) {
            const float x = tile[lane * P + j2];
            const float qv = sm.qs[c0 + j2];  // zero padded past d: (0, 0) leaves the accumulator as it is
            if (a.metric == 0) {
              acc = __fmaf_rn(qv, x, acc);
            } else {
              const float u = __fsub_rn(qv, x);
              acc = __fmaf_rn(u, u, acc);
            }
          }
        }
        if (have) atomicMax(&sm.best[sm.ent_slot[ei]], score_key(a.metric == 0 ? acc : -acc));
      }
    } else if ((a.d & 3) == 0) {
      const int d4 = a.d >> 2;
      const float4* qs4 = reinterpret_cast<const float4*>(sm.qs);
      for (int li = tid; li < nsurv; li += C::kThreads) {
        const int ei = (int)sm.surv[li];
        const float4* rp = reinterpret_cast<const float4*>(a.db_f32 + (size_t)sm.ent_row[ei] * a.d);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 cur[4], nxt[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) cur[t] = t < d4 ? __ldg(rp + t) : z;
        float acc = 0.0f;
        for (int k4 = 0; k4 < d4; k4 += 4) {
#pragma unroll
          for (int t = 0; t < 4; ++t) nxt[t] = k4 + 4 + t < d4 ? __ldg(rp + k4 + 4 + t) : z;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float4 qv = qs4[k4 + t];  // padded with zeros up to d_round
            if (a.metric == 0) {
              acc = __fmaf_rn(qv.x, cur[t].x, acc);
              acc = __fmaf_rn(qv.y, cur[t].y, acc);
              acc = __fmaf_rn(qv.z, cur[t].z, acc);
              acc = __fmaf_rn(qv.w, cur[t].w, acc);
            } else {  // padded columns: (0 - 0)^2
              float u = __fsub_rn(qv.x, cur[t].x); acc = __fmaf_rn(u, u, acc);
              u = __fsub_rn(qv.y, cur[t].y); acc = __fmaf_rn(u, u, acc);
              u = __fsub_rn(qv.z, cur[t].z); acc = __fmaf_rn(u, u, acc);
              u = __fsub_rn(qv.w, cur[t].w); acc = __fmaf_rn(u, u, acc);
            }
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) cur[t] = nxt[t];
        }
        atomicMax(&sm.best[sm.ent_slot[ei]], score_key(a.metric == 0 ? acc : -acc));
      }
    } else {  // rows are not 16-byte aligned: scalar walk
      for (int li = tid; li < nsurv; li += C::kThreads) {
        const int ei = (int)sm.surv[li];
        const float* rp = a.db_f32 + (size_t)sm.ent_row[ei] * a.d;
        float acc = 0.0f;
        for (int j = 0; j < a.d; ++j) {
          if (a.metric == 0) {
            acc = __fmaf_rn(sm.qs[j], __ldg(rp + j), acc);
          } else {
            const float u = __fsub_rn(sm.qs[j], __ldg(rp + j));
            acc = __fmaf_rn(u, u, acc);
          }
        }
        atomicMax(&sm.best[sm.ent_slot[ei]], score_key(a.metric == 0 ? acc : -acc));
      }
    }
  }

  // ---- E. one entry per session, best k of them (A overlays the record scratch: every pass is finished)
  __syncthreads();
  RF_PHASE(5);  // re-scoring
  // with k retained entries the old k-th key is a floor: every retained entry is >= it, a new session below it
  // cannot enter (a tie loses on the id) — dropping those here keeps the sort at ~k + newcomers elements
  const uint32_t floor_key = nr == a.k ? cand_key(g[a.k - 1]) : 0u;
  for (int h0 = warp * 32; h0 < C::kSlots; h0 += C::kThreads) {
    const int h = h0 + lane;
    const uint32_t o = sm.owner[h], bk = sm.best[h];
    const int pos = warp_append(&sm.ctr[2], o != 0u && bk != 0u && bk >= floor_key);
    if (pos >= 0) sm.A[pos] = pack_cand(bk, o - 1u);
  }
  __syncthreads();
  const int H = sm.ctr[2];
  int P = 2;
  while (P < H) P <<= 1;
  for (int i = H + tid; i < P; i += C::kThreads) sm.A[i] = 0ull;
  __syncthreads();
  sort_desc(sm.A, sm.tmp, H, P);
  const int m = H < a.k ? H : a.k;
  for (int i = tid; i < m; i += C::kThreads) g[i] = sm.A[i];
  if (tid == 0 && a.debug != nullptr) {  // volume counters for tuning (sss_index_stat 5..8)
    atomicAdd(&a.debug[0], (unsigned long long)n_ent);
    atomicAdd(&a.debug[1], (unsigned long long)sm.ctr[1]);
    atomicAdd(&a.debug[2], (unsigned long long)H);
    atomicAdd(&a.debug[3], 1ull);
  }
  RF_PHASE(6);  // final gather + sort + write back
  if (tid == 0) {
    st.cnt[q] = m;
    st.nret[q] = m;
    // thresholds only ever rise (a bootstrap threshold may already be in place while fewer than k sessions passed)
    st.thr[q] = fmaxf(thr_scan, to_scan(m == a.k ? fmaxf(new_thr, key_score(cand_key(sm.A[a.k - 1])) - margin) : new_thr));
  }
  return RF_DONE;
}

// one block per query; queries that do not fit go on the skip list of this wave
template <class C>
__global__ void __launch_bounds__(C::kThreads, 4) refine_small_kernel(RefineArgs a, SelectState st) {
  extern __shared__ __align__(16) unsigned char rf_smem[];
  RefineSmem<C> sm(rf_smem, (a.d + C::kKc - 1) / C::kKc * C::kKc);
  const int q = blockIdx.x;
  if (refine_query<C>(a, st, sm, q) == RF_SKIP && threadIdx.x == 0)
    st.skip_list[atomicAdd(&st.skip_cnt[a.wave & 1u], 1u)] = (uint32_t)q;
}

// persistent: drains the skip list (or every query when a.all_large)
template <class C>
__global__ void __launch_bounds__(C::kThreads) refine_large_kernel(RefineArgs a, SelectState st) {
  extern __shared__ __align__(16) unsigned char rf_smem[];
  RefineSmem<C> sm(rf_smem, (a.d + C::kKc - 1) / C::kKc * C::kKc);
  const uint32_t count = a.all_large ? (uint32_t)a.nq : st.skip_cnt[a.wave & 1u];
  if (blockIdx.x == 0 && threadIdx.x == 0) st.skip_cnt[(a.wave + 1u) & 1u] = 0u;  // for the next wave
  for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
    const int q = a.all_large ? (int)i : (int)st.skip_list[i];
    refine_query<C>(a, st, sm, q);
  }
}

// (128-thread blocks, 7 per SM — every query of a 1000-query wave resident at once — measured 1 % slower:
// profiles/r01_ab_experiments.md)
using RefineSmall = RefineCfg<10, 1536, 768, 512, 8, 16, 256, 512, false>;
using RefineLarge = RefineCfg<12, 4096, 4096, 2048, 16, 16, 512, 4096, true>;

int launch_refine(const RefineArgs& a_in, SelectState st, int num_sms, cudaStream_t stream) {
  RefineArgs a = a_in;
  SSS_REQUIRE(st.cap <= RefineLarge::kSlots, "candidate capacity too large for refine");
  SSS_REQUIRE(a.k <= RefineLarge::kKmax, "k too large for refine");
  SSS_REQUIRE(a.rec == nullptr || a.rec_nsub <= RefineLarge::kMaxSub, "too many record sub-regions per query");
  static SmemAttr attr_small, attr_large;  // per device
  a.all_large = a.k > RefineSmall::kKmax / 2 ? 1 : 0;
  if (!a.all_large) {
    const int d_round = (a.d + RefineSmall::kKc - 1) / RefineSmall::kKc * RefineSmall::kKc;
    const size_t smem = RefineSmall::smem_bytes(d_round, a.rescore != 0, a.d);
    SSS_REQUIRE(smem <= 226 * 1024, "embedding width too large for refine");
    if (attr_small.ensure(refine_small_kernel<RefineSmall>, (int)smem)) return 1;
    refine_small_kernel<RefineSmall><<<(unsigned)a.nq, RefineSmall::kThreads, smem, stream>>>(a, st);
    SSS_CUDA_OK(cudaGetLastError());
  }
  const int d_round = (a.d + RefineLarge::kKc - 1) / RefineLarge::kKc * RefineLarge::kKc;
  const size_t smem = RefineLarge::smem_bytes(d_round, a.rescore != 0, a.d);
  SSS_REQUIRE(smem <= 226 * 1024, "embedding width too large for refine");
  if (attr_large.ensure(refine_large_kernel<RefineLarge>, (int)smem)) return 1;
  int grid = (int)std::min<int64_t>(a.nq, num_sms);
  refine_large_kernel<RefineLarge><<<grid, RefineLarge::kThreads, smem, stream>>>(a, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- bootstrap thresholds ---------------------------------------------------------------------------------
// A tensor-core pass in chunk-max mode left cmax[chunk, q] = max score of 32 consecutive rows.  Among the
// (k-1) * gap + 1 largest chunk maxima of a query one can pick k chunks that are pairwise >= gap chunks apart
// (the kernel below looks at every gap-th chunk only, where any k chunks are),
// i.e. (with gap = floor((max session length + 30) / 32) + 1, or 1 without sessions) k rows of k DISTINCT
// sessions/rows whose scores are >= T, the smallest of those maxima.  So the k-th best exact session score is >= T - margin and a row
// can only matter if its tensor-core score is >= T - 2 * margin: thr = the float just below that.
// Only every gap-th chunk is looked at: any k of those are pairwise >= gap chunks apart, so the k-th largest of them
// is the threshold directly (the same quantile as the ((k-1) * gap + 1)-th largest of all chunks, from 1 / gap of the
// keys).  One block = 8 consecutive queries: the chunk maxima are stored [chunk][query], so 8 threads read one full
// 32-byte sector per chunk (a block per query read 4 of every 32 bytes it pulled: 128 MB of L2 traffic per 1000
// queries).  Warp w then finds the k-th largest key of query w by a bitwise binary search from the first bit in which
// the keys differ down to bit 8 — counting passes over shared memory, no atomics (a radix select's histogram
// serialises here: the maxima of one query share their leading bits, so a whole warp hits two or three bins).
constexpr int kBootQ = 8;
__global__ void __launch_bounds__(256) bootstrap_thr_kernel(const float* __restrict__ cmax, int n_chunks, int64_t nq,
                                                            int64_t nq_pad, int k, int chunk_gap, float slack_mult,
                                                            SelectState st) {
  extern __shared__ uint32_t bs_keys[];  // [kBootQ][n_s + 8] (row pitch keeps the transposing stores conflict free)
  const int n_s = n_chunks / chunk_gap;  // sampled chunks: 0, gap, 2 gap, ...
  const int pitch = n_s + 8;
  const int64_t q0 = (int64_t)blockIdx.x * kBootQ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int need = k;
  if (need > n_s) return;
  {
    // 16 independent loads in flight per thread: with a handful of blocks (few queries) this phase is pure latency
    const int qq = tid & 7;
    for (int i0 = tid >> 3; i0 < n_s; i0 += 32 * 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int i = i0 + 32 * u;
        v[u] = i < n_s ? cmax[(size_t)i * (size_t)chunk_gap * (size_t)nq_pad + (size_t)(q0 + qq)] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int i = i0 + 32 * u;
        if (i < n_s) bs_keys[qq * pitch + i] = score_key(v[u]);
      }
    }
  }
  __syncthreads();
  const int64_t q = q0 + warp;
  if (q >= nq) return;
  const uint32_t* keys = bs_keys + warp * pitch;
  // bits above the highest one in which the keys differ are common to all of them
  uint32_t diff = 0u;
  const uint32_t k0 = keys[0];
  for (int i = lane; i < n_s; i += 32) diff |= keys[i] ^ k0;
  diff = __reduce_or_sync(0xffffffffu, diff);
  const int top = diff ? 31 - __clz(diff) : -1;
  // largest v (low 8 bits left zero: any lower bound of the k-th largest key is a valid threshold, and 2^-15
  // relative is far inside the slack) with count(keys >= v) >= k, built bit by bit
  uint32_t prefix = top >= 31 ? 0u : (k0 & ~((2u << top) - 1u));
  if (top < 0) prefix = k0;
  for (int bit = top; bit >= 8; --bit) {
    const uint32_t cand = prefix | (1u << bit);
    int cnt = 0;
    int i = lane;
    for (; i + 96 < n_s; i += 128)
      cnt += (keys[i] >= cand ? 1 : 0) + (keys[i + 32] >= cand ? 1 : 0) + (keys[i + 64] >= cand ? 1 : 0) +
             (keys[i + 96] >= cand ? 1 : 0);
    for (; i < n_s; i += 32) cnt += keys[i] >= cand ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt >= need) prefix = cand;
  }
  if (lane == 0) st.thr[q] = nextafterf(key_score(prefix) - slack_mult * st.margin[q], -INFINITY);
}

int launch_bootstrap_thr(const float* cmax, int n_chunks, int64_t nq, int64_t nq_pad, int k, int chunk_gap,
                         float slack_mult, SelectState st, cudaStream_t stream) {
  SSS_REQUIRE(n_chunks <= 4096 && n_chunks % 32 == 0, "bootstrap region too large");
  static SmemAttr attr;
  const size_t smem = (size_t)kBootQ * (size_t)(n_chunks + 8) * 4;
  if (attr.ensure(bootstrap_thr_kernel, (int)smem)) return 1;
  bootstrap_thr_kernel<<<(unsigned)((nq + kBootQ - 1) / kBootQ), 256, smem, stream>>>(cmax, n_chunks, nq, nq_pad, k,
                                                                                     chunk_gap, slack_mult, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- emit -------------------------------------------------------------------------------------------
// status_out (optional): the search's status word (overflow bits | watchdog code << 8) is OR-ed into it, so that an
// asynchronous caller (the sharded path) can carry it inside the packed candidate block instead of stopping to read it
__global__ void emit_kernel(SelectState st, int64_t nq, int k, int metric, int64_t id_offset, float* __restrict__ D,
                            int64_t* __restrict__ I, int* __restrict__ status_out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 && status_out != nullptr) {
    const int v = st.overflow[0] | (st.overflow[1] << 8);
    if (v != 0) atomicOr(status_out, v);
  }
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

int launch_emit(SelectState st, int64_t nq, int k, int metric, int64_t id_offset, float* D, int64_t* I, int* status_out,
                cudaStream_t stream) {
  int64_t total = nq * k;
  if (total <= 0) return 0;
  emit_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(st, nq, k, metric, id_offset, D, I, status_out);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

const void* emit_kernel_addr() { return (const void*)emit_kernel; }

// ---- k-way merge of per-shard candidates (after the NCCL all-gather) ---------------------
// One block per query; (key desc, id asc) bitonic sort of n_shards*k (key, id64) pairs in shared memory.  Shard s
// holds its scores at cD + s * stride_d and its ids at cI + s * stride_i (elements): two [n_shards, nq, k] arrays, or
// the packed per-rank blocks [ids | scores] the sharded search all-gathers.
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ cD, const int64_t* __restrict__ cI,
                                                         int64_t stride_d, int64_t stride_i, int n_shards, int64_t nq,
                                                         int k, int metric, int P, float* __restrict__ D,
                                                         int64_t* __restrict__ I, const int* __restrict__ status_in,
                                                         int64_t status_stride, int* __restrict__ status_out) {
  extern __shared__ uint64_t sm[];
  if (blockIdx.x == 0 && threadIdx.x == 0 && status_out != nullptr) {  // OR of the shards' status words
    int v = 0;
    for (int s = 0; s < n_shards; ++s) v |= status_in[(int64_t)s * status_stride];
    *status_out = v;
  }
  uint64_t* keys = sm;                 // [P] key in the high word (0 = empty)
  int64_t* ids = (int64_t*)(sm + P);   // [P]
  const int64_t q = blockIdx.x;
  const int total = n_shards * k;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t kk = 0;
    int64_t id = INT64_MAX;
    if (i < total) {
      int s = i / k, j = i % k;
      int64_t gid = cI[(int64_t)s * stride_i + q * k + j];
      if (gid >= 0) {
        float sc = cD[(int64_t)s * stride_d + q * k + j];
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

int launch_topk_merge(const float* cD, const int64_t* cI, int64_t stride_d, int64_t stride_i, int n_shards, int64_t nq,
                      int k, int metric, float* D, int64_t* I, cudaStream_t stream, const int* status_in,
                      int64_t status_stride, int* status_out) {
  if (nq <= 0 || k <= 0) return 0;
  int total = n_shards * k;
  int P = 2;
  while (P < total) P <<= 1;
  size_t smem = (size_t)P * 16;
  SSS_REQUIRE(smem <= 128 * 1024, "sss_topk_merge: n_shards * k must be <= 8192");
  static SmemAttr attr;
  if (attr.ensure(topk_merge_kernel, (int)smem)) return 1;
  topk_merge_kernel<<<(unsigned)nq, 256, smem, stream>>>(cD, cI, stride_d, stride_i, n_shards, nq, k, metric, P, D, I,
                                                         status_in, status_stride, status_out);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
