// scan_bf16_sm100.cu — the fused tensor-core scan: Q x DB^T on tcgen05 (bf16 in, fp32 accumulate in
// TMEM), operands staged by TMA, and the streaming top-k FILTER fused into the epilogue so that the score
// matrix never leaves the SM.  Replaces the sgemm + heap inside faiss IndexFlatIP.search
// (test_amazon_filterd.py:211-214,578; fine_tune_ours.py:848-849,882).
//
// Four kernels share one structure (persistent, 1 CTA per SM, 384 threads, warp-specialised) — DESIGN.md 3.1:
//   scan_bf16_2cta_kernel   DEFAULT for d <= 128 and more than 128 queries: cluster (2,1,1), cta_group::2, M = 256
//                           queries x N = 256 rows per MMA, each CTA loads half of every 256-row DB tile
//   scan_bf16_ts_kernel     d <= 128, at most 128 queries (DB-stream bound): 1-CTA, query tiles live in TMEM
//   scan_bf16_kernel        d <= 128, 1-CTA, both operands from shared memory (kept for A/B runs: SSS_SCAN_VARIANT=ss)
//   scan_bf16_kloop_kernel  128 < d <= 4096: pair kernel with BOTH operands streamed per 64-wide K block
// Roles inside a CTA:
//   warp 0          TMA producer: resident query m-tiles once (or per K block), then DB tiles (128 rows x 64 columns
//                   of bf16, SWIZZLE_128B, 16 KB boxes) through an mbarrier ring
//   warp 1          MMA issuer (warp-uniform, one elected lane): tcgen05.mma kind::f16 into TMEM slots; commits to
//                   the slot's "full" barrier and, after a tile's last m-tile, to the stage's "empty" barrier
//   warp 2          TMEM allocator (512 columns)
//   warps 4-11      two epilogue warpgroups.  A thread owns ONE query (TMEM lane): tcgen05.ld 32 columns -> 3-input
//                   max tree -> one compare against the query's running threshold.  Only when some lane of the warp
//                   sees max > thr does the warp take the slow path: those lanes dump their 32 raw scores as a
//                   144-byte HitRecord into the private sub-region of (their query, this CTA or pair, their
//                   warpgroup) — a register counter, no atomics.  refine (select.cu) turns a query's records into
//                   exact top-k lists and a tighter threshold between waves.
//
// Arithmetic intensity: a CTA holding num_mt*128 resident queries does 2*num_mt*128*128*d_pad flop per
// 128*d_pad*2 bytes of DB tile, i.e. num_mt*128 flop/byte — 512 flop/B at num_mt=4, well above the B200
// ridge (~210-250 flop/B), so the kernel is tensor-bound for >= ~256 resident queries and HBM-bound (DB
// stream) below.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "kernels.h"
#include "tc_sm100.cuh"

namespace sss {

namespace {

struct ScanParams {
  int num_kb, num_mt, num_stages;
  int total_mtiles;      // over the whole (padded) query batch
  int64_t row_begin;     // multiple of 128
  int64_t n_tiles;       // DB tiles in this wave
  SelectState st;
  HitRecord* rec;
  uint32_t* rec_cnt;
  int rec_cap;
  int* err_flag;
  const uint4* q_bf16;   // [nq_pad, d_pad] bf16, 16-byte aligned rows (TS variant reads it directly)
  float* cmax;           // bootstrap mode: write the max of every 32-row chunk to cmax[chunk * nq_pad + q] instead
                         // of filtering (chunk counted from row_begin)
  int groups;            // K-loop variant: query groups of 256 (one per CTA pair)
  int last_k4;           // K-loop variant: 16-wide slices of real columns in the last K block (1..4)
};

template <int kCap>  // records per private sub-region (compile time: the epilogue is sensitive to it)
__global__ void __launch_bounds__(kNumThreads, 1)
scan_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                 const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int mt_base = blockIdx.y * p.num_mt;
  const int num_mt = min(p.num_mt, p.total_mtiles - mt_base);
  const int n_tiles = (int)p.n_tiles;
  const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  const uint32_t q_smem = smem_base;
  const uint32_t db_smem = q_smem + (uint32_t)(p.num_mt * p.num_kb) * kKBlockBytes;
  const uint32_t bar_base = db_smem + (uint32_t)(p.num_stages * p.num_kb) * kKBlockBytes;
  const uint32_t full_bar = bar_base;                         // [kMaxStages]
  const uint32_t empty_bar = bar_base + 8 * kMaxStages;       // [kMaxStages]
  const uint32_t tfull_bar = bar_base + 16 * kMaxStages;      // [4]
  const uint32_t tempty_bar = tfull_bar + 32;                 // [4]
  const uint32_t qfull_bar = tempty_bar + 32;                 // [1]
  const uint32_t tmem_ptr_addr = qfull_bar + 8;
  // generic pointer to the TMEM base slot for the post-alloc read
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_db);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 4);  // one arrive per epilogue warp of the owning warpgroup
    }
    mbar_init(qfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && my_tiles > 0) {
      mbar_expect_tx(qfull_bar, (uint32_t)(num_mt * p.num_kb) * kKBlockBytes);
      for (int mt = 0; mt < num_mt; ++mt)
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d(q_smem + (uint32_t)(mt * p.num_kb + kb) * kKBlockBytes, &tmap_q, qfull_bar, kb * 64,
                      (mt_base + mt) * kTileQ);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(empty_bar + 8 * stage, phase ^ 1u, p.err_flag, 101);
        mbar_expect_tx(full_bar + 8 * stage, (uint32_t)p.num_kb * kKBlockBytes);
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int row = (int)p.row_begin + tile * kTileRows;
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d(db_smem + (uint32_t)(stage * p.num_kb + kb) * kKBlockBytes, &tmap_db, full_bar + 8 * stage,
                      kb * 64, row);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (my_tiles > 0) {  // whole warp, uniform: one elected lane issues (keeps descriptors in uniform registers)
      mbar_wait(qfull_bar, 0, p.err_flag, 102);
      tc_fence_after();
      uint32_t u = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(full_bar + 8 * stage, phase, p.err_flag, 103);
        tc_fence_after();
        for (int mt = 0; mt < num_mt; ++mt, ++u) {
          const uint32_t slot = u & 3u;
          mbar_wait(tempty_bar + 8 * slot, ((u >> 2) & 1u) ^ 1u, p.err_flag, 104);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + slot * (uint32_t)kTileRows;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            const uint64_t adesc = umma_desc_sw128(q_smem + (uint32_t)(mt * p.num_kb + kb) * kKBlockBytes);
            const uint64_t bdesc = umma_desc_sw128(db_smem + (uint32_t)(stage * p.num_kb + kb) * kKBlockBytes);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)  // 4 x (K=16 bf16 = 32 bytes) inside the 128-byte swizzle row
              umma_bf16(d_tmem, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4), (kb | k4) != 0 ? 1u : 0u);
          }
          umma_commit(tfull_bar + 8 * slot);
        }
        umma_commit(empty_bar + 8 * stage);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: fused top-k filter =====================
    const int ew = warp - 4;
    const int wg = ew >> 2;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp may read
    // this thread's queries: m-tile mt -> query (mt_base + mt) * 128 + quarter * 32 + lane; one record counter each
    uint32_t rc0 = 0, rc1 = 0, rc2 = 0, rc3 = 0;
    const uint32_t sub_stride = gridDim.x * 2u;                          // sub-regions per query
    const uint32_t my_sub = blockIdx.x * 2u + (uint32_t)wg;
    const int total_units = my_tiles * num_mt;
    // unit u = (tile iteration it, resident m-tile mt); this warpgroup takes u = wg, wg+2, ...
    int it = 0, mt = wg;
    while (mt >= num_mt) { mt -= num_mt; ++it; }
    for (int u = wg; u < total_units; u += 2) {
      const uint32_t slot = (uint32_t)(u & 3);
      const uint32_t ph = (uint32_t)((u >> 2) & 1);
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int cur_mt = mt;
      const uint32_t qidx = (uint32_t)((mt_base + mt) * kTileQ + quarter * 32 + lane);
      const float thr = p.st.thr[qidx];
      const uint32_t row_tile = (uint32_t)((int)p.row_begin + tile * kTileRows);
      HitRecord* myrec = p.rec + ((size_t)qidx * sub_stride + my_sub) * (size_t)kCap;
      mt += 2;
      while (mt >= num_mt) { mt -= num_mt; ++it; }
      mbar_wait(tfull_bar + 8 * slot, ph, p.err_flag, 105);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * (uint32_t)kTileRows;
      // Two register buffers: the tcgen05.ld of chunk c+1 is in flight while chunk c goes through the max tree.
      auto process = [&](const uint32_t (&r)[32], int c) {
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
        float m0 = max3(f[0], f[1], f[2]), m1 = max3(f[3], f[4], f[5]);
        float m2 = max3(f[6], f[7], f[8]), m3 = max3(f[9], f[10], f[11]);
        m0 = max3(m0, f[12], f[13]); m1 = max3(m1, f[14], f[15]);
        m2 = max3(m2, f[16], f[17]); m3 = max3(m3, f[18], f[19]);
        m0 = max3(m0, f[20], f[21]); m1 = max3(m1, f[22], f[23]);
        m2 = max3(m2, f[24], f[25]); m3 = max3(m3, f[26], f[27]);
        m0 = max3(m0, f[28], f[29]); m1 = max3(m1, f[30], f[31]);
        const float mx = fmaxf(max3(m0, m1, m2), m3);
        const bool hit = mx > thr;
        if (__any_sync(0xffffffffu, hit)) {
          if (hit) {
            const uint32_t idx = cur_mt == 0 ? rc0 : cur_mt == 1 ? rc1 : cur_mt == 2 ? rc2 : rc3;
            if (idx < (uint32_t)kCap) {
              uint4* dst = reinterpret_cast<uint4*>(myrec + idx);
              dst[0] = make_uint4(qidx, row_tile + (uint32_t)(c * 32), 0u, 0u);
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[1 + i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
            }
            rc0 += cur_mt == 0; rc1 += cur_mt == 1; rc2 += cur_mt == 2; rc3 += cur_mt == 3;
          }
        }
      };
      uint32_t ra[32], rb[32];
      tmem_ld32(taddr, ra);
      tmem_ld_wait();
      tmem_ld32(taddr + 32u, rb);
      process(ra, 0);
      tmem_ld_wait();
      tmem_ld32(taddr + 64u, ra);
      process(rb, 1);
      tmem_ld_wait();
      tmem_ld32(taddr + 96u, rb);
      process(ra, 2);
      tmem_ld_wait();
      // all four chunks have left TMEM: hand the slot back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * slot);
      process(rb, 3);
    }
    // publish (and thereby reset) the record count of every sub-region this thread owns
    for (int m = 0; m < num_mt; ++m) {
      const uint32_t qidx = (uint32_t)((mt_base + m) * kTileQ + quarter * 32 + lane);
      p.rec_cnt[(size_t)qidx * sub_stride + my_sub] = m == 0 ? rc0 : m == 1 ? rc1 : m == 2 ? rc2 : rc3;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// TS variant: the resident query tiles live in TENSOR MEMORY (A operand from TMEM), so shared memory carries
// only the DB ring (7 stages at d=128) and a tcgen05.mma reads 4 KB instead of 8 KB of shared memory per K=16
// step — at N=128 the SS form sits exactly on the 128 B/cycle shared-memory port and the tensor pipe stalls.
// TMEM: [0, a_cols) = A (num_mt * num_kb * 32 columns: lane = query, column = two bf16), then two 128-column
// accumulator slots (slot == epilogue warpgroup).
// ---------------------------------------------------------------------------------------------------------
template <int kCap, bool kFp8>  // records per private sub-region (compile time: the epilogue is sensitive to it);
                                // kFp8: operands are E4M3 bytes (Hamming search over +-1 codes), same byte geometry
__global__ void __launch_bounds__(kNumThreads, 1)
scan_bf16_ts_kernel(const __grid_constant__ CUtensorMap tmap_db, const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int mt_base = blockIdx.y * p.num_mt;
  const int num_mt = min(p.num_mt, p.total_mtiles - mt_base);
  const int n_tiles = (int)p.n_tiles;
  const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  const uint32_t db_smem = smem_base;
  const uint32_t bar_base = db_smem + (uint32_t)(p.num_stages * p.num_kb) * kKBlockBytes;
  const uint32_t full_bar = bar_base;                         // [kMaxStages]
  const uint32_t empty_bar = bar_base + 8 * kMaxStages;       // [kMaxStages]
  const uint32_t tfull_bar = bar_base + 16 * kMaxStages;      // [2]
  const uint32_t tempty_bar = tfull_bar + 16;                 // [2]
  const uint32_t qready_bar = tempty_bar + 16;                // [1]
  const uint32_t tmem_ptr_addr = qready_bar + 8;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_db);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 4);  // one arrive per epilogue warp of the owning warpgroup
    }
    mbar_init(qready_bar, 8);            // every epilogue warp has stored its share of Q
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;
  const uint32_t a_cols = (uint32_t)(p.num_mt * p.num_kb) * 32u;
  const uint32_t d_base = tmem_base + a_cols;

  if (warp == 0) {
    // ===================== TMA producer: DB tiles only =====================
    if (lane == 0 && my_tiles > 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(empty_bar + 8 * stage, phase ^ 1u, p.err_flag, 201);
        mbar_expect_tx(full_bar + 8 * stage, (uint32_t)p.num_kb * kKBlockBytes);
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int row = (int)p.row_begin + tile * kTileRows;
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d(db_smem + (uint32_t)(stage * p.num_kb + kb) * kKBlockBytes, &tmap_db, full_bar + 8 * stage,
                      kb * 64, row);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (my_tiles > 0) {  // whole warp, uniform: one elected lane issues
      mbar_wait(qready_bar, 0, p.err_flag, 202);
      tc_fence_after();
      uint32_t u = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(full_bar + 8 * stage, phase, p.err_flag, 203);
        tc_fence_after();
        for (int mt = 0; mt < num_mt; ++mt, ++u) {
          const uint32_t slot = u & 1u;
          mbar_wait(tempty_bar + 8 * slot, ((u >> 1) & 1u) ^ 1u, p.err_flag, 204);
          tc_fence_after();
          const uint32_t d_tmem = d_base + slot * (uint32_t)kTileRows;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            const uint32_t a_tmem = tmem_base + (uint32_t)(mt * p.num_kb + kb) * 32u;
            const uint64_t bdesc = umma_desc_sw128(db_smem + (uint32_t)(stage * p.num_kb + kb) * kKBlockBytes);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {  // K=16 bf16 (32 fp8) = 8 TMEM columns of A, 32 bytes of the B swizzle row
              if (kFp8)
                umma_fp8_ts(d_tmem, a_tmem + (uint32_t)(8 * k4), bdesc + (uint64_t)(2 * k4), (kb | k4) != 0 ? 1u : 0u);
              else
                umma_bf16_ts(d_tmem, a_tmem + (uint32_t)(8 * k4), bdesc + (uint64_t)(2 * k4), (kb | k4) != 0 ? 1u : 0u);
            }
          }
          umma_commit(tfull_bar + 8 * slot);
        }
        umma_commit(empty_bar + 8 * stage);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int ew = warp - 4;
    const int wg = ew >> 2;
    const int quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    // ---- stage this warp's share of the query tiles into TMEM (warpgroup wg takes m-tiles wg, wg+2)
    if (my_tiles > 0) {
      const int row_words = p.num_kb * 8;  // uint4 per query row
      for (int mt = wg; mt < num_mt; mt += 2) {
        const uint4* src = p.q_bf16 + (size_t)((mt_base + mt) * kTileQ + quarter * 32 + lane) * row_words;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = src[kb * 8 + i];
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
          }
          tmem_st32(tmem_base + lane_base + (uint32_t)(mt * p.num_kb + kb) * 32u, w);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qready_bar);
    }
    // ---- fused top-k filter
    uint32_t rc0 = 0, rc1 = 0, rc2 = 0, rc3 = 0;
    const uint32_t sub_stride = gridDim.x * 2u;
    const uint32_t my_sub = blockIdx.x * 2u + (uint32_t)wg;
    const int total_units = my_tiles * num_mt;
    int it = 0, mt = wg;
    while (mt >= num_mt) { mt -= num_mt; ++it; }
    for (int u = wg; u < total_units; u += 2) {
      const uint32_t slot = (uint32_t)wg;  // u & 1
      const uint32_t ph = (uint32_t)((u >> 1) & 1);
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int cur_mt = mt;
      const uint32_t qidx = (uint32_t)((mt_base + mt) * kTileQ + quarter * 32 + lane);
      const float thr = p.st.thr[qidx];
      const uint32_t row_tile = (uint32_t)((int)p.row_begin + tile * kTileRows);
      HitRecord* myrec = p.rec + ((size_t)qidx * sub_stride + my_sub) * (size_t)kCap;
      mt += 2;
      while (mt >= num_mt) { mt -= num_mt; ++it; }
      mbar_wait(tfull_bar + 8 * slot, ph, p.err_flag, 205);
      tc_fence_after();
      const uint32_t taddr = d_base + lane_base + slot * (uint32_t)kTileRows;
      auto process = [&](const uint32_t (&r)[32], int c) {
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
        float m0 = max3(f[0], f[1], f[2]), m1 = max3(f[3], f[4], f[5]);
        float m2 = max3(f[6], f[7], f[8]), m3 = max3(f[9], f[10], f[11]);
        m0 = max3(m0, f[12], f[13]); m1 = max3(m1, f[14], f[15]);
        m2 = max3(m2, f[16], f[17]); m3 = max3(m3, f[18], f[19]);
        m0 = max3(m0, f[20], f[21]); m1 = max3(m1, f[22], f[23]);
        m2 = max3(m2, f[24], f[25]); m3 = max3(m3, f[26], f[27]);
        m0 = max3(m0, f[28], f[29]); m1 = max3(m1, f[30], f[31]);
        const float mx = fmaxf(max3(m0, m1, m2), m3);
        if (p.cmax != nullptr) {  // bootstrap pass: chunk maxima only (coalesced: lane = query)
          const uint32_t chunk = (row_tile - (uint32_t)p.row_begin) / 32u + (uint32_t)c;
          p.cmax[(size_t)chunk * (size_t)(p.total_mtiles * kTileQ) + qidx] = mx;
          return;
        }
        const bool hit = mx > thr;
        if (__any_sync(0xffffffffu, hit)) {
          if (hit) {
            const uint32_t idx = cur_mt == 0 ? rc0 : cur_mt == 1 ? rc1 : cur_mt == 2 ? rc2 : rc3;
            if (idx < (uint32_t)kCap) {
              uint4* dst = reinterpret_cast<uint4*>(myrec + idx);
              dst[0] = make_uint4(qidx, row_tile + (uint32_t)(c * 32), 0u, 0u);
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[1 + i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
            }
            rc0 += cur_mt == 0; rc1 += cur_mt == 1; rc2 += cur_mt == 2; rc3 += cur_mt == 3;
          }
        }
      };
      uint32_t ra[32], rb[32];
      tmem_ld32(taddr, ra);
      tmem_ld_wait();
      tmem_ld32(taddr + 32u, rb);
      process(ra, 0);
      tmem_ld_wait();
      tmem_ld32(taddr + 64u, ra);
      process(rb, 1);
      tmem_ld_wait();
      tmem_ld32(taddr + 96u, rb);
      process(ra, 2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * slot);
      process(rb, 3);
    }
    for (int m = 0; m < num_mt; ++m) {
      const uint32_t qidx = (uint32_t)((mt_base + m) * kTileQ + quarter * 32 + lane);
      p.rec_cnt[(size_t)qidx * sub_stride + my_sub] = m == 0 ? rc0 : m == 1 ? rc1 : m == 2 ? rc2 : rc3;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// 2-CTA variant (thread-block cluster of two CTAs on one TPC, tcgen05 cta_group::2).  One MMA covers M = 256
// queries (128 from each CTA's shared memory) x N = 256 DB rows (128 from each CTA's shared memory): every CTA
// loads only HALF of each DB tile, the single issuing thread (leader CTA) dispatches instructions that are
// twice as long (128 cycles), and both tensor cores run off one instruction stream.  Each CTA keeps its own
// queries' accumulators in its own TMEM (2 slots x 256 columns) and runs its own epilogue.
//   full[stage]    lives in the leader; both CTAs' TMA loads complete_tx on it (cta_group::2 TMA form)
//   empty[stage]   one per CTA, signalled by a multicast tcgen05.commit from the leader
//   tfull[slot]    one per CTA, multicast commit
//   tempty[slot]   leader only, 8 arrivals: the 4 epilogue warps of the owning warpgroup in BOTH CTAs
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc2), "r"(accumulate)
      : "memory");
}
constexpr uint32_t kIdesc2Fp8 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
__device__ __forceinline__ void umma_fp8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc2Fp8), "r"(accumulate)
      : "memory");
}
template <int kCap, bool kFp8>  // records per private sub-region (compile time: the epilogue is sensitive to it)
__global__ void __launch_bounds__(kNumThreads, 1)
scan_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                      const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = (int)(blockIdx.x >> 1);
  const int n_pairs = (int)(gridDim.x >> 1);
  const int nq_pad = p.total_mtiles * kTileQ;

  // m-tiles of this CTA: group base g0, unit j -> m-tile g0 + 2 * j + rank
  const int g0 = blockIdx.y * 2 * p.num_mt;
  const int mt_in_group = min(2 * p.num_mt, p.total_mtiles - g0);
  const int num_j = (mt_in_group + 1) / 2;                   // MMA units per DB tile (same in both CTAs)
  const int n_tiles = (int)p.n_tiles;                        // 256-row tiles in this wave
  const int my_tiles = pair < n_tiles ? (n_tiles - 1 - pair) / n_pairs + 1 : 0;

  const uint32_t q_smem = smem_base;
  const uint32_t db_smem = q_smem + (uint32_t)(p.num_mt * p.num_kb) * kKBlockBytes;
  const uint32_t bar_base = db_smem + (uint32_t)(p.num_stages * p.num_kb) * kKBlockBytes;
  const uint32_t full_bar = bar_base;                         // [kMaxStages] (leader's are used)
  const uint32_t empty_bar = bar_base + 8 * kMaxStages;       // [kMaxStages]
  const uint32_t tfull_bar = bar_base + 16 * kMaxStages;      // [2]
  const uint32_t tempty_bar = tfull_bar + 16;                 // [2] (leader's are used)
  const uint32_t qfull_bar = tempty_bar + 16;                 // [1]
  const uint32_t tmem_ptr_addr = qfull_bar + 8;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_db);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 8);
    }
    mbar_init(qfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  // resident query tiles of this CTA; the leader may only start once BOTH CTAs hold theirs
  if (warp == 0 && lane == 0 && my_tiles > 0) {
    mbar_expect_tx(qfull_bar, (uint32_t)(num_j * p.num_kb) * kKBlockBytes);
    for (int j = 0; j < num_j; ++j)
      for (int kb = 0; kb < p.num_kb; ++kb)
        tma_load_2d(q_smem + (uint32_t)(j * p.num_kb + kb) * kKBlockBytes, &tmap_q, qfull_bar, kb * 64,
                    (g0 + 2 * j + (int)rank) * kTileQ);
    mbar_wait(qfull_bar, 0, p.err_flag, 302);
  }
  cluster_sync_all();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own half of every DB tile =====================
    if (lane == 0 && my_tiles > 0) {
      const uint32_t leader_full = map_to_cta(full_bar, 0);  // the leader's barrier, shared::cluster address
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(empty_bar + 8 * stage, phase ^ 1u, p.err_flag, 301);
        if (leader) mbar_expect_tx(full_bar + 8 * stage, 2u * (uint32_t)p.num_kb * kKBlockBytes);
        const int tile = pair + it * n_pairs;
        const int row = (int)p.row_begin + tile * 256 + (int)rank * kTileRows;
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d_2sm(db_smem + (uint32_t)(stage * p.num_kb + kb) * kKBlockBytes, &tmap_db,
                          leader_full + 8 * stage, kb * 64, row);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && my_tiles > 0) {  // whole warp, uniform: one elected lane issues
      tc_fence_after();
      uint32_t u = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(full_bar + 8 * stage, phase, p.err_flag, 303);
        tc_fence_after();
        for (int j = 0; j < num_j; ++j, ++u) {
          const uint32_t slot = u & 1u;
          mbar_wait(tempty_bar + 8 * slot, ((u >> 1) & 1u) ^ 1u, p.err_flag, 304);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + slot * 256u;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            const uint64_t adesc = umma_desc_sw128(q_smem + (uint32_t)(j * p.num_kb + kb) * kKBlockBytes);
            const uint64_t bdesc = umma_desc_sw128(db_smem + (uint32_t)(stage * p.num_kb + kb) * kKBlockBytes);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              if (kFp8)
                umma_fp8_2cta(d_tmem, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4), (kb | k4) != 0 ? 1u : 0u);
              else
                umma_bf16_2cta(d_tmem, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4), (kb | k4) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2cta(tfull_bar + 8 * slot);
        }
        umma_commit_2cta(empty_bar + 8 * stage);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): fused top-k filter over 256 columns per unit ==========
    const int ew = warp - 4;
    const int wg = ew >> 2;
    const int quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t leader_tempty = map_to_cta(tempty_bar, 0);
    uint32_t rc0 = 0, rc1 = 0, rc2 = 0, rc3 = 0;
    const uint32_t sub_stride = (uint32_t)n_pairs * 2u;
    const uint32_t my_sub = (uint32_t)pair * 2u + (uint32_t)wg;
    const int total_units = my_tiles * num_j;
    int it = 0, j = wg;
    while (j >= num_j) { j -= num_j; ++it; }
    for (int u = wg; u < total_units; u += 2) {
      const uint32_t slot = (uint32_t)wg;
      const uint32_t ph = (uint32_t)((u >> 1) & 1);
      const int tile = pair + it * n_pairs;
      const int cur_j = j;
      const uint32_t qidx = (uint32_t)((g0 + 2 * j + (int)rank) * kTileQ + quarter * 32 + lane);
      const bool q_ok = qidx < (uint32_t)nq_pad;
      const float thr = q_ok ? p.st.thr[qidx] : INFINITY;
      const uint32_t row_tile = (uint32_t)((int)p.row_begin + tile * 256);
      HitRecord* myrec = p.rec + ((size_t)(q_ok ? qidx : 0u) * sub_stride + my_sub) * (size_t)kCap;
      j += 2;
      while (j >= num_j) { j -= num_j; ++it; }
      mbar_wait(tfull_bar + 8 * slot, ph, p.err_flag, 305);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + slot * 256u;
      auto process = [&](const uint32_t (&r)[32], int c) {
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
        float m0 = max3(f[0], f[1], f[2]), m1 = max3(f[3], f[4], f[5]);
        float m2 = max3(f[6], f[7], f[8]), m3 = max3(f[9], f[10], f[11]);
        m0 = max3(m0, f[12], f[13]); m1 = max3(m1, f[14], f[15]);
        m2 = max3(m2, f[16], f[17]); m3 = max3(m3, f[18], f[19]);
        m0 = max3(m0, f[20], f[21]); m1 = max3(m1, f[22], f[23]);
        m2 = max3(m2, f[24], f[25]); m3 = max3(m3, f[26], f[27]);
        m0 = max3(m0, f[28], f[29]); m1 = max3(m1, f[30], f[31]);
        const float mx = fmaxf(max3(m0, m1, m2), m3);
        if (p.cmax != nullptr) {
          if (q_ok) {
            const uint32_t chunk = (row_tile - (uint32_t)p.row_begin) / 32u + (uint32_t)c;
            p.cmax[(size_t)chunk * (size_t)nq_pad + qidx] = mx;
          }
          return;
        }
        const bool hit = mx > thr;
        if (__any_sync(0xffffffffu, hit)) {
          if (hit) {
            const uint32_t idx = cur_j == 0 ? rc0 : cur_j == 1 ? rc1 : cur_j == 2 ? rc2 : rc3;
            if (idx < (uint32_t)kCap) {
              uint4* dst = reinterpret_cast<uint4*>(myrec + idx);
              dst[0] = make_uint4(qidx, row_tile + (uint32_t)(c * 32), 0u, 0u);
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[1 + i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
            }
            rc0 += cur_j == 0; rc1 += cur_j == 1; rc2 += cur_j == 2; rc3 += cur_j == 3;
          }
        }
      };
      uint32_t ra[32], rb[32];
      tmem_ld32(taddr, ra);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        tmem_ld32(taddr + (uint32_t)((c + 1) * 32), rb);
        process(ra, c);
        tmem_ld_wait();
        if (c + 2 < 8) {
          tmem_ld32(taddr + (uint32_t)((c + 2) * 32), ra);
        } else {  // all eight chunks have left TMEM: hand the slot back to the leader's MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar + 8 * slot);
            else mbar_arrive_remote(leader_tempty + 8 * slot);
          }
        }
        process(rb, c + 1);
        if (c + 2 < 8) tmem_ld_wait();
      }
    }
    for (int m = 0; m < num_j; ++m) {
      const uint32_t qidx = (uint32_t)((g0 + 2 * m + (int)rank) * kTileQ + quarter * 32 + lane);
      if (qidx < (uint32_t)nq_pad)
        p.rec_cnt[(size_t)qidx * sub_stride + my_sub] = m == 0 ? rc0 : m == 1 ? rc1 : m == 2 ? rc2 : rc3;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ======================================================================================================
// scan_bf16_kloop_kernel — any width (d_pad a multiple of 64, up to 4096): BOTH operands are streamed per
// 64-wide K block, the accumulators stay in TMEM for the whole K sweep.  This is the tensor path of the
// encoder-shaped embeddings (d = 1600, model/gnn.py:184-191), whose query tiles do not fit in shared memory.
//
// Cluster (2,1,1), cta_group::2.  Pair p serves ONE query group g = p % G (256 queries: 128 per CTA) and the DB
// super-tiles T = p / G + n * (pairs / G); a super-tile is 512 rows = two 256-row MMA tiles, one per 256-column
// TMEM slot, so that one query K block feeds eight MMAs:
//   stage = [A: 128 queries x 64 | B0: my 128 rows of tile 2T x 64 | B1: my 128 rows of tile 2T+1 x 64] = 48 KB per CTA
//   per stage 2 x 4 MMAs of M=256 x N=256 x K=16 = 1024 tensor cycles per 48 KB per SM (47 B/clk, under the
//   ~42-64 B/clk/SM an SM can pull from L2; a single 256-row tile per sweep would need 64 B/clk).
// Pairs p .. p+G-1 walk the same super-tile at the same time with different query groups, so a DB tile comes from
// HBM once and from L2 for the other groups; the query matrix (nq_pad x d_pad bf16, a few MB) lives in L2.
// Both slots are drained after the sweep (warpgroup w takes slot w) — no TMEM double buffering: the exposed
// epilogue is ~2000 cycles against num_kb * 1024 cycles of MMA (7 % at d = 1600).
constexpr int kKloopStageBytes = 3 * kKBlockBytes;

__global__ void __launch_bounds__(kNumThreads, 1)
scan_bf16_kloop_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                       const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = (int)(blockIdx.x >> 1);
  const int n_pairs = (int)(gridDim.x >> 1);
  const int G = p.groups;
  const int g = pair % G;            // query group of this pair
  const int pig = pair / G;          // pair index inside the group
  const int ppg = n_pairs / G;       // pairs per group (the grid is G * ppg pairs)
  const int nq_pad = p.total_mtiles * kTileQ;
  const int n_super = (int)p.n_tiles;  // 512-row super-tiles in this wave
  const int my_items = pig < n_super ? (n_super - 1 - pig) / ppg + 1 : 0;
  const int mt = 2 * g + (int)rank;    // m-tile of this CTA (may be one past the end for an odd tile count)

  const uint32_t bar_base = smem_base + (uint32_t)(p.num_stages * kKloopStageBytes);
  const uint32_t full_bar = bar_base;                         // [kMaxStages] (leader's are used)
  const uint32_t empty_bar = bar_base + 8 * kMaxStages;       // [kMaxStages]
  const uint32_t tfull_bar = bar_base + 16 * kMaxStages;      // [1]
  const uint32_t tempty_bar = tfull_bar + 8;                  // [1] (leader's is used)
  const uint32_t tmem_ptr_addr = tempty_bar + 8;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_db);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 16);  // 8 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own half of A, B0, B1 for every K block =====================
    if (lane == 0 && my_items > 0) {
      const uint32_t leader_full = map_to_cta(full_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_items; ++it) {
        const int T = pig + it * ppg;
        const int row0 = (int)p.row_begin + T * 512 + (int)rank * kTileRows;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1u, p.err_flag, 401);
          if (leader) mbar_expect_tx(full_bar + 8 * stage, 2u * (uint32_t)kKloopStageBytes);
          const uint32_t dst = smem_base + (uint32_t)(stage * kKloopStageBytes);
          tma_load_2d_2sm(dst, &tmap_q, leader_full + 8 * stage, kb * 64, mt * kTileQ);
          tma_load_2d_2sm(dst + kKBlockBytes, &tmap_db, leader_full + 8 * stage, kb * 64, row0);
          tma_load_2d_2sm(dst + 2 * kKBlockBytes, &tmap_db, leader_full + 8 * stage, kb * 64, row0 + 256);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && my_items > 0) {  // whole warp, uniform: one elected lane issues
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_items; ++it) {
        mbar_wait(tempty_bar, ((uint32_t)it & 1u) ^ 1u, p.err_flag, 404);  // both slots drained (item it - 1)
        tc_fence_after();
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase, p.err_flag, 403);
          tc_fence_after();
          const uint32_t sbase = smem_base + (uint32_t)(stage * kKloopStageBytes);
          const uint64_t adesc = umma_desc_sw128(sbase);
          const uint64_t b0desc = umma_desc_sw128(sbase + kKBlockBytes);
          const uint64_t b1desc = umma_desc_sw128(sbase + 2 * kKBlockBytes);
          // the last K block may hold fewer than four 16-wide slices of real columns (the rest is zero padding)
          const int n4 = kb + 1 == p.num_kb ? p.last_k4 : 4;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            if (k4 < n4) {
              const uint32_t acc = (kb | k4) != 0 ? 1u : 0u;
              umma_bf16_2cta(tmem_base, adesc + (uint64_t)(2 * k4), b0desc + (uint64_t)(2 * k4), acc);
              umma_bf16_2cta(tmem_base + 256u, adesc + (uint64_t)(2 * k4), b1desc + (uint64_t)(2 * k4), acc);
            }
          }
          umma_commit_2cta(empty_bar + 8 * stage);
          if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2cta(tfull_bar);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): warpgroup w drains slot w (256 columns) ==========
    const int wg = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t leader_tempty = map_to_cta(tempty_bar, 0);
    const uint32_t qidx = (uint32_t)(mt * kTileQ + quarter * 32 + lane);
    const bool q_ok = qidx < (uint32_t)nq_pad;
    const uint32_t sub_stride = (uint32_t)ppg * 2u;
    const uint32_t my_sub = (uint32_t)pig * 2u + (uint32_t)wg;
    HitRecord* myrec = p.rec + ((size_t)(q_ok ? qidx : 0u) * sub_stride + my_sub) * (size_t)p.rec_cap;
    uint32_t rc = 0;
    for (int it = 0; it < my_items; ++it) {
      const int T = pig + it * ppg;
      const uint32_t row_tile = (uint32_t)((int)p.row_begin + T * 512 + wg * 256);
      const float thr = q_ok ? p.st.thr[qidx] : INFINITY;
      mbar_wait(tfull_bar, (uint32_t)it & 1u, p.err_flag, 405);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + (uint32_t)wg * 256u;
      auto process = [&](const uint32_t (&r)[32], int c) {
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
        float m0 = max3(f[0], f[1], f[2]), m1 = max3(f[3], f[4], f[5]);
        float m2 = max3(f[6], f[7], f[8]), m3 = max3(f[9], f[10], f[11]);
        m0 = max3(m0, f[12], f[13]); m1 = max3(m1, f[14], f[15]);
        m2 = max3(m2, f[16], f[17]); m3 = max3(m3, f[18], f[19]);
        m0 = max3(m0, f[20], f[21]); m1 = max3(m1, f[22], f[23]);
        m2 = max3(m2, f[24], f[25]); m3 = max3(m3, f[26], f[27]);
        m0 = max3(m0, f[28], f[29]); m1 = max3(m1, f[30], f[31]);
        const float mx = fmaxf(max3(m0, m1, m2), m3);
        if (p.cmax != nullptr) {
          if (q_ok) {
            const uint32_t chunk = (row_tile - (uint32_t)p.row_begin) / 32u + (uint32_t)c;
            p.cmax[(size_t)chunk * (size_t)nq_pad + qidx] = mx;
          }
          return;
        }
        const bool hit = mx > thr;
        if (__any_sync(0xffffffffu, hit)) {
          if (hit) {
            if (rc < (uint32_t)p.rec_cap) {
              uint4* dst = reinterpret_cast<uint4*>(myrec + rc);
              dst[0] = make_uint4(qidx, row_tile + (uint32_t)(c * 32), 0u, 0u);
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[1 + i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
            }
            ++rc;
          }
        }
      };
      uint32_t ra[32], rb[32];
      tmem_ld32(taddr, ra);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        tmem_ld32(taddr + (uint32_t)((c + 1) * 32), rb);
        process(ra, c);
        tmem_ld_wait();
        if (c + 2 < 8) {
          tmem_ld32(taddr + (uint32_t)((c + 2) * 32), ra);
        } else {  // all eight chunks have left TMEM: hand the slot back to the leader's MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar);
            else mbar_arrive_remote(leader_tempty);
          }
        }
        process(rb, c + 1);
        if (c + 2 < 8) tmem_ld_wait();
      }
    }
    if (q_ok) p.rec_cnt[(size_t)qidx * sub_stride + my_sub] = rc;
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace

int make_tensor_map_bf16_2d(void* out_map128, const void* base, uint64_t rows, uint64_t cols_pad, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  SSS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  SSS_REQUIRE(cols_pad % 64 == 0, "bf16 operand width must be padded to a multiple of 64");
  SSS_REQUIRE(((uintptr_t)base & 127) == 0, "bf16 operand base must be 128-byte aligned");
  cuuint64_t dims[2] = {cols_pad, rows};
  cuuint64_t strides[1] = {cols_pad * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn((CUtensorMap*)out_map128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

int plan_scan_bf16(int d_pad, int64_t nq_pad, int num_sms, int max_stages, Bf16ScanPlan* plan, int rec_boost,
                   int d_used, int variant) {
  SSS_REQUIRE(d_pad % 64 == 0 && d_pad >= 64 && d_pad <= 4096, "tensor-core scan supports d <= 4096");
  SSS_REQUIRE(nq_pad % kTileQ == 0 && nq_pad > 0, "nq_pad must be a positive multiple of 128");
  const int total_mtiles = (int)(nq_pad / kTileQ);
  plan->kloop = false;
  plan->fp8 = false;
  plan->groups = 0;
  {
    const int tail = (d_used > 0 ? d_used : d_pad) - (d_pad - 64);  // real columns of the last K block
    plan->last_k4 = tail <= 0 ? 4 : (tail + 15) / 16 > 4 ? 4 : (tail + 15) / 16;
  }
  if (d_pad > 128) {
    // wide rows: the K-loop pair kernel (both operands streamed per K block)
    const int G = (total_mtiles + 1) / 2;
    const int n_pairs = num_sms / 2;
    SSS_REQUIRE(G <= n_pairs, "query batch too large for the K-loop scan (split it)");
    const int ppg = n_pairs / G;
    plan->kloop = true;
    plan->ts = false;
    plan->two_cta = false;
    plan->groups = G;
    plan->num_kb = d_pad / 64;
    plan->num_mt = 1;
    plan->total_mtiles = total_mtiles;
    plan->grid_x = 2 * G * ppg;
    plan->grid_y = 1;
    int stages = (227 * 1024 - 1024 - kBarrierBytes) / kKloopStageBytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > max_stages) stages = max_stages;
    SSS_REQUIRE(stages >= 2, "not enough shared memory for the operand ring");
    plan->num_stages = stages;
    plan->smem_bytes = 1024 + stages * kKloopStageBytes + kBarrierBytes;
    plan->rec_cap = 4 * kRecSubCap * rec_boost;  // a query has ppg * 2 sub-regions here, a quarter of the d <= 128 kernels'
    plan->rec_nsub = ppg * 2;
    plan->n_regions = (int)nq_pad * plan->rec_nsub;
    plan->tile_rows = 512;
    return 0;
  }
  plan->num_kb = d_pad / 64;
  plan->total_mtiles = total_mtiles;
  plan->grid_y = (total_mtiles + 3) / 4;
  plan->num_mt = (total_mtiles + plan->grid_y - 1) / plan->grid_y;
  plan->grid_x = num_sms / plan->grid_y;
  if (plan->grid_x < 1) plan->grid_x = 1;
  // variant (a tuning knob read once per handle, SSS_SCAN_VARIANT): 0 automatic, 1 SS, 2 TS, 3 pair
  plan->ts = variant != 1;  // A operand from TMEM unless the SS form is forced
  plan->two_cta = variant == 0 || variant == 3;
  // one m-tile (<= 128 queries) is the DB-stream-bound regime: independent 1-CTA tiles with the queries in TMEM
  // reach 0.86 of the HBM copy bandwidth, the pair kernel (half of every MMA is padding there) 0.77
  if (total_mtiles == 1 && variant != 3) plan->two_cta = false;
  if (plan->two_cta) {
    // pairs of CTAs: every pair holds up to 8 m-tiles (4 per CTA) and walks 256-row DB tiles
    plan->ts = false;
    plan->num_kb = d_pad / 64;
    plan->total_mtiles = total_mtiles;
    plan->grid_y = (total_mtiles + 7) / 8;
    const int per_group = (total_mtiles + plan->grid_y - 1) / plan->grid_y;  // m-tiles per pair
    plan->num_mt = (per_group + 1) / 2;                                      // per CTA
    int pairs = (num_sms / 2) / plan->grid_y;
    if (pairs < 1) pairs = 1;
    plan->grid_x = 2 * pairs;
    const int q_bytes2 = plan->num_mt * plan->num_kb * kKBlockBytes;
    const int stage_bytes2 = plan->num_kb * kKBlockBytes;
    int stages2 = (227 * 1024 - 1024 - kBarrierBytes - q_bytes2) / stage_bytes2;
    if (stages2 > kMaxStages) stages2 = kMaxStages;
    if (stages2 > max_stages) stages2 = max_stages;
    SSS_REQUIRE(stages2 >= 2, "not enough shared memory for the DB tile ring");
    plan->num_stages = stages2;
    plan->smem_bytes = 1024 + q_bytes2 + stages2 * stage_bytes2 + kBarrierBytes;
    plan->rec_cap = kRecSubCap * rec_boost;
    plan->rec_nsub = pairs * 2;
    plan->n_regions = (int)nq_pad * plan->rec_nsub;
    plan->tile_rows = 256;
    return 0;
  }
  plan->tile_rows = kTileRows;
  const int q_bytes = plan->ts ? 0 : plan->num_mt * plan->num_kb * kKBlockBytes;
  const int stage_bytes = plan->num_kb * kKBlockBytes;
  const int avail = 227 * 1024 - 1024 - kBarrierBytes - q_bytes;
  int stages = avail / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > max_stages) stages = max_stages;
  SSS_REQUIRE(stages >= 2, "not enough shared memory for the DB tile ring");
  plan->num_stages = stages;
  plan->smem_bytes = 1024 + q_bytes + stages * stage_bytes + kBarrierBytes;
  plan->rec_cap = kRecSubCap * rec_boost;
  plan->rec_nsub = plan->grid_x * 2;
  plan->n_regions = (int)nq_pad * plan->rec_nsub;
  return 0;
}

int launch_scan_bf16(const Bf16ScanPlan& plan, const void* tmap_q, const void* tmap_db, const void* q_bf16,
                     int64_t row_begin, int64_t row_end, SelectState st, HitRecord* rec, uint32_t* rec_cnt,
                     int* err_flag, float* cmax, cudaStream_t stream) {
  SSS_REQUIRE(cmax == nullptr || plan.ts || plan.two_cta || plan.kloop,
              "chunk-max bootstrap needs the TS, 2-CTA or K-loop scan variant");
  SSS_REQUIRE(row_begin % plan.tile_rows == 0, "scan wave must start on a tile boundary");
  ScanParams p;
  p.num_kb = plan.num_kb;
  p.num_mt = plan.num_mt;
  p.num_stages = plan.num_stages;
  p.total_mtiles = plan.total_mtiles;
  p.row_begin = row_begin;
  p.n_tiles = (row_end - row_begin + plan.tile_rows - 1) / plan.tile_rows;
  p.st = st;
  p.rec = rec;
  p.rec_cnt = rec_cnt;
  p.rec_cap = plan.rec_cap;
  p.err_flag = err_flag;
  p.q_bf16 = (const uint4*)q_bf16;
  p.cmax = cmax;
  p.groups = plan.groups;
  p.last_k4 = plan.last_k4;
  if (p.n_tiles <= 0) return 0;
  const bool big = plan.rec_cap != kRecSubCap && !plan.kloop;  // boosted sub-regions (after an overflow): 4x
  SSS_REQUIRE(plan.kloop || plan.rec_cap == kRecSubCap || plan.rec_cap == 4 * kRecSubCap, "unsupported record capacity");
  // (per device, per kernel: the attribute lives in the device's context)
  static SmemAttr a_ss, a_ss4, a_ts, a_ts4, a_pair, a_pair4, a_kloop, a_ts8, a_ts84, a_pair8, a_pair84;
  const int sb = plan.smem_bytes;
  SSS_REQUIRE(!plan.fp8 || (!plan.kloop && (plan.two_cta || plan.ts)), "the fp8 (Hamming) scan runs on the pair and TS kernels");
  if (plan.kloop) {
    if (a_kloop.ensure(scan_bf16_kloop_kernel, sb)) return 1;
  } else if (plan.two_cta && plan.fp8) {
    if (big ? a_pair84.ensure(scan_bf16_2cta_kernel<4 * kRecSubCap, true>, sb) : a_pair8.ensure(scan_bf16_2cta_kernel<kRecSubCap, true>, sb)) return 1;
  } else if (plan.two_cta) {
    if (big ? a_pair4.ensure(scan_bf16_2cta_kernel<4 * kRecSubCap, false>, sb) : a_pair.ensure(scan_bf16_2cta_kernel<kRecSubCap, false>, sb)) return 1;
  } else if (plan.ts && plan.fp8) {
    if (big ? a_ts84.ensure(scan_bf16_ts_kernel<4 * kRecSubCap, true>, sb) : a_ts8.ensure(scan_bf16_ts_kernel<kRecSubCap, true>, sb)) return 1;
  } else if (plan.ts) {
    if (big ? a_ts4.ensure(scan_bf16_ts_kernel<4 * kRecSubCap, false>, sb) : a_ts.ensure(scan_bf16_ts_kernel<kRecSubCap, false>, sb)) return 1;
  } else {
    if (big ? a_ss4.ensure(scan_bf16_kernel<4 * kRecSubCap>, sb) : a_ss.ensure(scan_bf16_kernel<kRecSubCap>, sb)) return 1;
  }
  dim3 grid(plan.grid_x, plan.grid_y);
  const CUtensorMap& mq = *(const CUtensorMap*)tmap_q;
  const CUtensorMap& mdb = *(const CUtensorMap*)tmap_db;
  if (plan.two_cta || plan.kloop) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (plan.kloop)
      SSS_CUDA_OK(cudaLaunchKernelEx(&cfg, scan_bf16_kloop_kernel, mq, mdb, p));
    else if (plan.fp8 && big)
      SSS_CUDA_OK(cudaLaunchKernelEx(&cfg, scan_bf16_2cta_kernel<4 * kRecSubCap, true>, mq, mdb, p));
    else if (plan.fp8)
      SSS_CUDA_OK(cudaLaunchKernelEx(&cfg, scan_bf16_2cta_kernel<kRecSubCap, true>, mq, mdb, p));
    else if (big)
      SSS_CUDA_OK(cudaLaunchKernelEx(&cfg, scan_bf16_2cta_kernel<4 * kRecSubCap, false>, mq, mdb, p));
    else
      SSS_CUDA_OK(cudaLaunchKernelEx(&cfg, scan_bf16_2cta_kernel<kRecSubCap, false>, mq, mdb, p));
  } else if (plan.ts) {
    if (plan.fp8 && big)
      scan_bf16_ts_kernel<4 * kRecSubCap, true><<<grid, kNumThreads, plan.smem_bytes, stream>>>(mdb, p);
    else if (plan.fp8)
      scan_bf16_ts_kernel<kRecSubCap, true><<<grid, kNumThreads, plan.smem_bytes, stream>>>(mdb, p);
    else if (big)
      scan_bf16_ts_kernel<4 * kRecSubCap, false><<<grid, kNumThreads, plan.smem_bytes, stream>>>(mdb, p);
    else
      scan_bf16_ts_kernel<kRecSubCap, false><<<grid, kNumThreads, plan.smem_bytes, stream>>>(mdb, p);
  } else {
    if (big)
      scan_bf16_kernel<4 * kRecSubCap><<<grid, kNumThreads, plan.smem_bytes, stream>>>(mq, mdb, p);
    else
      scan_bf16_kernel<kRecSubCap><<<grid, kNumThreads, plan.smem_bytes, stream>>>(mq, mdb, p);
  }
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
