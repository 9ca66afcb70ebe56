// gemm_bf16x3_sm100.cu — the encoder's dense linears on the tensor cores with fp32-level accuracy, with the
// element-wise work that follows every linear in the reference fused into the epilogue.
//
// C[M, N] = A[M, K] * B[N, K]^T with both operands split into two bf16 terms, x = hi + lo (hi = bf16(x),
// lo = bf16(x - hi): 16 mantissa bits together), and three tcgen05 products accumulated in fp32 in TMEM:
//   A*B ~= Ah*Bh + Ah*Bl + Al*Bh      (the dropped Al*Bl term and the split residuals are ~2^-16 relative).
// Through the whole encoder (three HeteroGGNN layers + pooling, 768 -> 3 x 800 -> 3168 -> 1600) this stays 2.7e-5 of
// the output scale from a float64 forward, ten times inside the parity tolerance.  Replaces the x @ W.T of
// torch.nn.Linear / GATConv.lin_src / GRUCell inside model/gnn.py:64-81,193-217.
//
// Persistent CTAs (one per SM, 384 threads, warp-specialised) walk the 128 x BN output tiles (BN = 128, or 96 for the
// GRU GEMM) of up to two independent problems — the query-side and the product-side linears of a layer, or the two
// pooling projections, run as ONE launch:
//   warp 0  TMA producer: per 64-wide K block one stage [Ah | Al | Bh | Bl] of SWIZZLE_128B boxes through a 3-stage
//           ring; A is read at a column offset (the layer's slice of the concatenated node embeddings), the last K
//           block may hold fewer than four 16-wide slices of real columns
//   warp 1  MMA issuer: 3 products x 4 (K = 16) tcgen05.mma kind::f16 per stage into one of TWO TMEM accumulators, so
//           the next tile's main loop runs while the previous tile is still in its epilogue
//   warp 2  TMEM allocator
//   warps 4-11  epilogue, two warpgroups: both walk the accumulator 32 columns at a time and take 16 of them each
//           (a warp may only read the TMEM lanes of its quarter, so the split is by columns): tcgen05.ld (a thread owns
//           one output row; row-wise reductions finish here), then a [128][17] shared-memory transpose per warpgroup so
//           that 16 consecutive threads touch 16 consecutive columns of global memory (a thread-per-row epilogue wrote
//           32 different sectors per instruction: the GRU epilogue alone took 90 us per launch that way)
// What bounds it at the encoder's sizes (profiles/r02_encoder.md, A/B switches in a scratch build): with one or two
// tiles per SM per launch the LAST epilogue of every CTA is exposed, and with one epilogue warpgroup it was a quarter
// of the forward (0.59 ms -> 0.44 ms with the phase-2 work removed; one MMA product instead of three changed nothing).
// The second warpgroup halves it: 0.557 -> 0.445 ms per 200-session forward.  Neither form of cta_group::2 pays at
// these sizes: 256 x 256 pair tiles (half the operand bytes per flop) 0.78 vs 0.69 ms, 256 x BN pair tiles (three
// quarters of the bytes, same tile count, four stages) 0.460 vs 0.445 ms — the cross-CTA hop per stage costs more
// than the L2 traffic it saves; an L2 prefetch of the next launch's weights by the idle warp changed nothing either.
//
// Epilogues (what the reference does right after the linear, fused so that no activation makes an extra trip
// through HBM and no separate split / rowdot / GRU / tanh kernel is launched):
//   EPI_STORE    C (+ bias) as fp32; optionally sign() (BinarizeHead, model/model.py:137)
//   EPI_ATT      C as fp32 + per-row partial dot products with the GAT attention vectors (GATConv's a_s / a_d,
//                SURVEY appendix A), one partial per (row, N tile) summed in a fixed order by the consumer
//   EPI_GRU      GRUCell gates + HeteroConv sum + relu (model/gnn.py:59,72): the weight rows are permuted so that
//                an N tile of 96 columns holds the r, z, n gates of the same 32 hidden units; writes the next layer's
//                product features as fp32 AND as the hi / lo bf16 operand of the next GEMM
//   EPI_POOL     tanh([lin + b | PE[pos]]) of PositionalAttentionPooling (model/gnn.py:199-206) including the
//                repeat_interleave of product rows by their occurrence count; writes U as fp32 and hi / lo
//   EPI_ATTPOOL  partial sums of w_att . sigmoid(node_emb_lin(U) + b + coarse_rep_lin(c)[graph]) (model/gnn.py:213-215)
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_sm100.cuh"

namespace sss {

namespace {

constexpr int kGemmThreads = 384;             // 4 role warps + 2 epilogue warpgroups
constexpr int kGemmStageBytes = 4 * kKBlockBytes;  // Ah | Al | Bh | Bl
constexpr int kGemmStages = 3;
constexpr uint32_t kGemmTmemCols = 256;            // two accumulators of 128 columns
constexpr int kGemmXposePitch = 17;
constexpr int kGemmXposeBytes = 2 * 128 * kGemmXposePitch * 4;  // one [128][17] transpose buffer per epilogue warpgroup
constexpr int kGemmSmemBytes = 1024 + kGemmStages * kGemmStageBytes + kGemmXposeBytes + 256;

struct GemmLaunch {
  GemmProblem p[2];
  int n_problems;
  int tiles0;       // tiles of problem 0 (global tile index below this belongs to it)
  int tiles_total;
  int* err_flag;
};

__device__ __forceinline__ void umma_bf16_n(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate,
                                                 uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void store_hilo(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}
// sigmoid / tanh on the special-function unit (ex2.approx + rcp.approx, ~1e-6 absolute): the epilogue warps are the
// only ones doing arithmetic on their scheduler, and the libm forms (25 instructions for tanhf) made the epilogue of
// the GRU and pooling GEMMs longer than their main loops
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  const float xc = fminf(fmaxf(x, -15.0f), 15.0f);
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * xc));
}
// the 4 warps of epilogue warpgroup wg (named barriers 1 and 2)
__device__ __forceinline__ void epi_bar(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

struct TileRef {
  int which, tile_m, tile_n, m0, n0;
};
__device__ __forceinline__ TileRef tile_of(const GemmLaunch& L, int t) {
  TileRef r;
  r.which = t >= L.tiles0 ? 1 : 0;
  const GemmProblem& g = L.p[r.which];
  const int local = t - (r.which ? L.tiles0 : 0);
  r.tile_n = local % g.tiles_n;
  r.tile_m = local / g.tiles_n;
  r.m0 = r.tile_m * kTileQ;
  r.n0 = r.tile_n * g.bn;
  return r;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_ah0, const __grid_constant__ CUtensorMap tm_al0,
                   const __grid_constant__ CUtensorMap tm_bh0, const __grid_constant__ CUtensorMap tm_bl0,
                   const __grid_constant__ CUtensorMap tm_ah1, const __grid_constant__ CUtensorMap tm_al1,
                   const __grid_constant__ CUtensorMap tm_bh1, const __grid_constant__ CUtensorMap tm_bl1,
                   const __grid_constant__ GemmLaunch L) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = (int)blockIdx.x;       // (tile walker: this CTA, stride = grid)
  const int n_pairs = (int)gridDim.x;

  const uint32_t xpose_smem = smem_base + (uint32_t)(kGemmStages * kGemmStageBytes);
  const uint32_t bar_base = xpose_smem + (uint32_t)kGemmXposeBytes;
  const uint32_t full_bar = bar_base;             // [kGemmStages]
  const uint32_t empty_bar = bar_base + 32;       // [kGemmStages]
  const uint32_t tfull_bar = bar_base + 64;       // [2]
  const uint32_t tempty_bar = bar_base + 80;      // [2]
  const uint32_t tmem_ptr_addr = bar_base + 96;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
  float* const xpose = reinterpret_cast<float*>(smem_raw + (xpose_smem - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_ah0);
    tma_prefetch_desc(&tm_al0);
    tma_prefetch_desc(&tm_bh0);
    tma_prefetch_desc(&tm_bl0);
    if (L.n_problems > 1) {
      tma_prefetch_desc(&tm_ah1);
      tma_prefetch_desc(&tm_al1);
      tma_prefetch_desc(&tm_bh1);
      tma_prefetch_desc(&tm_bl1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 8);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr),
                 "r"(kGemmTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < L.tiles_total; t += n_pairs) {
        const TileRef tr = tile_of(L, t);
        const GemmProblem& g = L.p[tr.which];
        const CUtensorMap* ah = tr.which ? &tm_ah1 : &tm_ah0;
        const CUtensorMap* al = tr.which ? &tm_al1 : &tm_al0;
        const CUtensorMap* bh = tr.which ? &tm_bh1 : &tm_bh0;
        const CUtensorMap* bl = tr.which ? &tm_bl1 : &tm_bl0;
        const uint32_t b_bytes = (uint32_t)g.bn * 128u;  // one B box: bn rows x 64 bf16
        const int my_m = tr.m0, my_n = tr.n0;
        for (int kb = 0; kb < g.num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1u, L.err_flag, 501);
          mbar_expect_tx(full_bar + 8 * stage, 2u * (uint32_t)kKBlockBytes + 2u * b_bytes);
          const uint32_t dst = smem_base + (uint32_t)(stage * kGemmStageBytes);
          tma_load_2d(dst, ah, full_bar + 8 * stage, g.a_k0 + kb * 64, my_m);
          tma_load_2d(dst + kKBlockBytes, al, full_bar + 8 * stage, g.a_k0 + kb * 64, my_m);
          tma_load_2d(dst + 2 * kKBlockBytes, bh, full_bar + 8 * stage, kb * 64, my_n);
          tma_load_2d(dst + 3 * kKBlockBytes, bl, full_bar + 8 * stage, kb * 64, my_n);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) ======================
    {
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int t = pair; t < L.tiles_total; t += n_pairs, ++it) {
        const TileRef tr = tile_of(L, t);
        const GemmProblem& g = L.p[tr.which];
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.bn >> 3) << 17) | ((uint32_t)(kTileQ >> 4) << 24);
        const uint32_t slot = it & 1u;
        mbar_wait(tempty_bar + 8 * slot, ((it >> 1) & 1u) ^ 1u, L.err_flag, 504);  // the accumulator was drained
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + slot * 128u;
        for (int kb = 0; kb < g.num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase, L.err_flag, 503);
          tc_fence_after();
          const uint32_t sbase = smem_base + (uint32_t)(stage * kGemmStageBytes);
          const uint64_t ah = umma_desc_sw128(sbase);
          const uint64_t al = umma_desc_sw128(sbase + kKBlockBytes);
          const uint64_t bh = umma_desc_sw128(sbase + 2 * kKBlockBytes);
          const uint64_t bl = umma_desc_sw128(sbase + 3 * kKBlockBytes);
          // the last K block may hold fewer than four 16-wide slices of real columns: what follows them in the A
          // buffer belongs to the next layer's slice and must not enter the product
          const int n4 = kb + 1 == g.num_kb ? g.last_k4 : 4;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            if (k4 < n4) {
              const uint64_t o = (uint64_t)(2 * k4);
              // small terms first, the dominant product last (all of them accumulate in fp32 anyway)
              umma_bf16_n(d_tmem, al + o, bh + o, (kb | k4) != 0 ? 1u : 0u, idesc);
              umma_bf16_n(d_tmem, ah + o, bl + o, 1u, idesc);
              umma_bf16_n(d_tmem, ah + o, bh + o, 1u, idesc);
            }
          }
          umma_commit(empty_bar + 8 * stage);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar + 8 * slot);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps = 2 warpgroups) =====================
    // Both warpgroups read every 32-column chunk of the accumulator, 16 columns each (a warp may only touch the TMEM
    // lanes of its quarter, so the tile is split by columns, not rows): TMEM -> registers (a thread owns one output
    // row; row-wise reductions finish here) -> [128][17] shared-memory transpose -> phase 2, where 16 consecutive
    // threads touch 16 consecutive columns of one row of global memory.
    const int ew = warp - 4;
    const int wg = ew >> 2;
    const int quarter = warp & 3;
    const int rl = quarter * 32 + lane;                  // phase 1: this thread's row inside the tile
    const int e = (int)threadIdx.x - 128 - wg * 128;     // phase 2: e & 15 = column inside the half chunk, e >> 4 = row group
    const int pc = e & 15, pr = e >> 4;
    const int hc = 16 * wg;                              // this warpgroup's columns inside a chunk
    float* const xp = xpose + wg * (kTileQ * kGemmXposePitch);
    uint32_t it = 0;
    for (int t = pair; t < L.tiles_total; t += n_pairs, ++it) {
      const TileRef tr = tile_of(L, t);
      const GemmProblem& g = L.p[tr.which];
      const int m0 = tr.m0, n0 = tr.n0, tile_n = tr.tile_n;
      const uint32_t slot = it & 1u;
      const int row = m0 + rl;
      const bool row_ok = row < g.M;
      const int rows_valid = min(kTileQ, g.M - m0);
      mbar_wait(tfull_bar + 8 * slot, (it >> 1) & 1u, L.err_flag, 505);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * 128u + (uint32_t)hc;
      const int n_chunks = g.bn / 32;
      float partial = 0.0f;
      const float* bc = (g.epi == EPI_ATTPOOL && row_ok) ? g.ap_bc + (size_t)g.ap_node_graph[row] * (size_t)g.N : nullptr;
      float rg[16], zg[16];  // EPI_GRU: gates of this thread's phase-2 elements (row group k, unit pc)
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (c + 1 == n_chunks) {  // the accumulator has left TMEM: hand the slot back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar + 8 * slot);
        }
        if (g.epi == EPI_ATTPOOL) {
          if (row_ok) {
            const int col0 = n0 + c * 32 + hc;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = col0 + i;
              if (col < g.N) partial += g.ap_w[col] * fast_sigmoid(__uint_as_float(r[i]) + g.bias[col] + bc[col]);
            }
          }
          continue;
        }
        if (g.epi == EPI_ATT) {
          // attention logits: one partial per (row, 128-column sub-tile, warpgroup); a sub-tile lies inside one part
          const int sub = (n0 + c * 32) / 128;                 // 128-column sub-tile of the whole output
          const int part = sub / g.att_tiles_per_part;
          const float* att = part < 4 ? g.att[part] : nullptr;
          if (att != nullptr) {
            const int pc0 = (sub - part * g.att_tiles_per_part) * 128 + (c & 3) * 32 + hc;  // column inside the part
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (pc0 + i < g.att_width) partial = fmaf(__uint_as_float(r[i]), att[pc0 + i], partial);
            if ((c & 3) == 3) {
              if (row_ok)
                g.att_out[part][((size_t)row * g.att_tiles_per_part + (sub - part * g.att_tiles_per_part)) * 2 + wg] = partial;
              partial = 0.0f;
            }
          }
        }
        epi_bar(wg);  // the previous chunk's readers are done with the transpose buffer
#pragma unroll
        for (int i = 0; i < 16; ++i) xp[rl * kGemmXposePitch + i] = __uint_as_float(r[i]);
        epi_bar(wg);
        // ---- phase 2: element (row group k -> row 8k + pr, column pc of this half chunk)
        if (g.epi == EPI_STORE || g.epi == EPI_ATT) {
          const int col = n0 + c * 32 + hc + pc;
          if (col < g.N) {
            const float bias = g.bias ? g.bias[col] : 0.0f;
            for (int k = 0; 8 * k + pr < rows_valid; ++k) {
              const int rr = 8 * k + pr;
              float v = xp[rr * kGemmXposePitch + pc] + bias;
              if (g.sign_out) {
                const float sg = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);
                const float th = tanhf(v);
                v = (sg - th) + th;  // (sign - tanh).detach() + tanh, model/model.py:137
              }
              g.C[(size_t)(m0 + rr) * (size_t)g.ldc + col] = v;
            }
          }
        } else if (g.epi == EPI_GRU) {
          // every 96 columns = 32 hidden units x (r | z | n): chunk c -> unit group c / 3, gate c % 3
          const int H = g.gru_H;
          const int gate = c % 3;
          const int u = (tile_n * (g.bn / 96) + c / 3) * 32 + hc + pc;
          if (u < H) {
            const float bi = g.gru_b_ih[gate * H + u], bh = g.gru_b_hh[gate * H + u];
            // all global operands of the chunk are requested before the first one is used: independent loads in
            // flight instead of serialised L2 round trips (160 us per launch that way)
            const float* ghp = g.gru_gh + (size_t)(m0 + pr) * (size_t)g.gru_gh_ld + gate * H + u;
            const size_t gh_step = 8 * (size_t)g.gru_gh_ld;
            float ghv[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) ghv[k] = 8 * k + pr < rows_valid ? __ldg(ghp + k * gh_step) : 0.0f;
            if (gate < 2) {
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float gi = xp[(8 * k + pr) * kGemmXposePitch + pc] + bi;
                const float sgm = fast_sigmoid(gi + ghv[k] + bh);
                if (gate == 0) rg[k] = sgm; else zg[k] = sgm;
              }
            } else {
              float xv[16], gpv[16];
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const int rr = 8 * k + pr;
                const bool ok = rr < rows_valid;
                xv[k] = (ok && u < g.gru_in_w) ? __ldg(g.gru_x + (size_t)(m0 + rr) * (size_t)g.gru_x_ld + u) : 0.0f;
                gpv[k] = ok ? __ldg(g.gru_gp + (size_t)(m0 + rr) * (size_t)H + u) : 0.0f;
              }
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const int rr = 8 * k + pr;
                if (rr < rows_valid) {
                  const float gi = xp[rr * kGemmXposePitch + pc] + bi;
                  const float ng = fast_tanh(gi + rg[k] * (ghv[k] + bh));
                  const float h = (1.0f - zg[k]) * ng + zg[k] * xv[k];
                  const float o = fmaxf(gpv[k] + h, 0.0f);  // HeteroConv sum of both branches, relu
                  g.C[(size_t)(m0 + rr) * (size_t)g.ldc + u] = o;
                  store_hilo(g.out_hi, g.out_lo, (size_t)(m0 + rr) * g.out_ld + g.out_k0 + u, o);
                }
              }
            }
          }
        } else if (g.epi == EPI_POOL) {
          const int W = g.pool_lin_w + g.pool_msl;
          const int col = n0 + c * 32 + hc + pc;
          if (col < W) {
            const bool is_lin = col < g.pool_lin_w;
            const float bias = is_lin ? g.bias[col] : 0.0f;
            // output rows of an input row: a query -> one row after the product occurrences; a product -> one per
            // occurrence (repeat_interleave by cnt).  Row ranges first (independent loads), then the stores.
            int t0v[16], t1v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const int rr = 8 * k + pr;
              const int grow = m0 + rr;
              t0v[k] = 0;
              t1v[k] = 0;
              if (rr < rows_valid) {
                if (g.pool_is_product) {
                  t0v[k] = __ldg(g.pool_prefix + grow);
                  t1v[k] = __ldg(g.pool_prefix + grow + 1);
                } else {
                  t0v[k] = g.pool_row0 + grow;
                  t1v[k] = t0v[k] + 1;
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const int rr = 8 * k + pr;
              const float lin = is_lin ? fast_tanh(xp[rr * kGemmXposePitch + pc] + bias) : 0.0f;
              for (int tt = t0v[k]; tt < t1v[k]; ++tt) {
                float x = lin;
                if (!is_lin) {
                  const int64_t pos = g.pool_pos[g.pool_is_product ? tt : m0 + rr];
                  x = fast_tanh(g.pool_pe[pos * g.pool_msl + (col - g.pool_lin_w)]);
                }
                g.pool_U[(size_t)tt * W + col] = x;
                store_hilo(g.out_hi, g.out_lo, (size_t)tt * g.out_ld + col, x);
              }
            }
          }
        }
      }
      if (row_ok && g.epi == EPI_ATTPOOL) g.ap_out[((size_t)row * g.tiles_n + tile_n) * 2 + wg] = partial;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kGemmTmemCols)
                 : "memory");
  }
}

// x[rows, cols] fp32 (row stride ld; element (r, c) at x[r * ld + c], or x[c * ld + r] when transposed) ->
// hi, lo bf16 [rows_pad, cols_pad] zero padded; with row_map, output row r takes source row row_map[r] (< 0: zeros)
__global__ void split_bf16_kernel(const float* __restrict__ x, int rows, int cols, int64_t ld, int transposed,
                                  const int* __restrict__ row_map, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int rows_pad, int cols_pad) {
  const int64_t total = (int64_t)rows_pad * cols_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ro = (int)(i / cols_pad), c = (int)(i % cols_pad);
    const int r = row_map ? row_map[ro] : ro;
    float v = 0.0f;
    if (r >= 0 && r < rows && c < cols) v = transposed ? x[(int64_t)c * ld + r] : x[(int64_t)r * ld + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace

int launch_split_bf16(const float* x, int rows, int cols, int64_t ld, int transposed, void* hi, void* lo, int rows_pad,
                      int cols_pad, cudaStream_t stream, const int* row_map) {
  // (rows_pad is the number of rows WRITTEN, zero beyond `rows`: a whole operand, or one part of a fused weight)
  SSS_REQUIRE(cols_pad % 64 == 0 && (row_map != nullptr || rows_pad >= rows) && cols_pad >= cols, "split_bf16: bad padded shape");
  const int64_t total = (int64_t)rows_pad * cols_pad;
  if (total == 0) return 0;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)current_sm_count() * 16);
  split_bf16_kernel<<<blocks, 256, 0, stream>>>(x, rows, cols, ld, transposed, row_map, (__nv_bfloat16*)hi,
                                                (__nv_bfloat16*)lo, rows_pad, cols_pad);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gemm_bf16x3(const GemmProblem* problems, int n_problems, int* err_flag, cudaStream_t stream) {
  SSS_REQUIRE(n_problems == 1 || n_problems == 2, "gemm_bf16x3: one or two problems per launch");
  GemmLaunch L;
  alignas(64) unsigned char tm[8][128];
  int tiles[2] = {0, 0};
  for (int i = 0; i < 2; ++i) {
    const GemmProblem& g = problems[i < n_problems ? i : 0];
    SSS_REQUIRE(g.bn == 128 || g.bn == 96, "gemm_bf16x3: N tile must be 128 or 96");
    SSS_REQUIRE(g.num_kb >= 1 && g.last_k4 >= 1 && g.last_k4 <= 4 && g.tiles_n >= 1, "gemm_bf16x3: bad K / N tiling");
    SSS_REQUIRE(g.a_rows_pad % 128 == 0 && g.a_ld % 64 == 0 && g.b_ld % 64 == 0, "gemm_bf16x3: bad operand pitch");
    SSS_REQUIRE((g.a_k0 * 2) % 16 == 0, "gemm_bf16x3: A column offset must be 16-byte aligned");
    SSS_REQUIRE(g.b_rows_pad >= g.tiles_n * g.bn, "gemm_bf16x3: the weight operand must be padded to whole tiles");
    if (make_tensor_map_bf16_2d(tm[4 * i + 0], g.a_hi, (uint64_t)g.a_rows_pad, (uint64_t)g.a_ld, 128)) return 1;
    if (make_tensor_map_bf16_2d(tm[4 * i + 1], g.a_lo, (uint64_t)g.a_rows_pad, (uint64_t)g.a_ld, 128)) return 1;
    if (make_tensor_map_bf16_2d(tm[4 * i + 2], g.b_hi, (uint64_t)g.b_rows_pad, (uint64_t)g.b_ld, (uint32_t)g.bn)) return 1;
    if (make_tensor_map_bf16_2d(tm[4 * i + 3], g.b_lo, (uint64_t)g.b_rows_pad, (uint64_t)g.b_ld, (uint32_t)g.bn)) return 1;
    L.p[i] = g;
    if (i < n_problems) tiles[i] = ((g.M + 127) / 128) * g.tiles_n;
  }
  L.n_problems = n_problems;
  L.tiles0 = tiles[0];
  L.err_flag = err_flag;
  const int total = tiles[0] + tiles[1];
  L.tiles_total = total;
  if (total <= 0) return 0;
  static SmemAttr attr;  // per device
  if (attr.ensure(gemm_bf16x3_kernel, kGemmSmemBytes)) return 1;
  const int n_sm = current_sm_count();
  const int grid = total < n_sm ? total : n_sm;  // persistent: one CTA per SM walks the tiles
  gemm_bf16x3_kernel<<<(unsigned)grid, kGemmThreads, kGemmSmemBytes, stream>>>(
      *(const CUtensorMap*)tm[0], *(const CUtensorMap*)tm[1], *(const CUtensorMap*)tm[2], *(const CUtensorMap*)tm[3],
      *(const CUtensorMap*)tm[4], *(const CUtensorMap*)tm[5], *(const CUtensorMap*)tm[6], *(const CUtensorMap*)tm[7], L);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
