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
// One CTA per 128 x BN output tile (BN = 128, or 96 for the GRU GEMM), 256 threads, warp-specialised:
//   warp 0  TMA producer: per 64-wide K block one stage [Ah | Al | Bh | Bl] of SWIZZLE_128B boxes; A is read at a
//           column offset (the layer's slice of the concatenated node embeddings), the last K block may hold
//           fewer than four 16-wide slices of real columns
//   warp 1  MMA issuer: 3 products x 4 (K = 16) tcgen05.mma kind::f16 per stage into one accumulator
//   warp 2  TMEM allocator
//   warps 4-7  epilogue: a thread owns one output row; 32 accumulator columns at a time
// A launch carries up to two independent problems (blockIdx.x enumerates the tiles of both): the query-side and the
// product-side linears of a layer, or the two pooling projections, run as ONE launch.
//
// Epilogues (what the reference does right after the linear, fused so that no activation makes an extra trip
// through HBM and no separate split / rowdot / GRU / tanh kernel is launched):
//   EPI_STORE    C (+ bias) as fp32; optionally sign() (BinarizeHead, model/model.py:137)
//   EPI_ATT      C as fp32 + per-row partial dot products with the GAT attention vectors (GATConv's a_s / a_d,
//                SURVEY appendix A), one partial per (row, N tile) summed in a fixed order by the consumer
//   EPI_GRU      GRUCell gates + HeteroConv sum + relu (model/gnn.py:59,72): the weight rows are permuted so that an
//                N tile of 96 columns holds the r, z, n gates of the same 32 hidden units; writes the next layer's
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

constexpr int kGemmThreads = 256;
constexpr int kGemmStageBytes = 4 * kKBlockBytes;  // Ah | Al | Bh | Bl
constexpr int kGemmStages = 3;
constexpr uint32_t kGemmTmemCols = 128;
constexpr int kGemmSmemBytes = 1024 + kGemmStages * kGemmStageBytes + 256;

struct GemmLaunch {
  GemmProblem p[2];
  int n_problems;
  int tiles0;  // tiles of problem 0 (blockIdx.x below this belongs to it)
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

// ---- epilogues: thread = one output row, r[32] = accumulator columns [col0, col0 + 32) of the tile --------------
__device__ __forceinline__ void epi_store(const GemmProblem& g, int row, int n0, int c, const uint32_t (&r)[32]) {
  float* crow = g.C + (size_t)row * (size_t)g.ldc;
  const int col0 = n0 + c * 32;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int col = col0 + i;
    if (col < g.N) {
      float v = __uint_as_float(r[i]);
      if (g.bias) v += g.bias[col];
      if (g.sign_out) {
        const float s = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);
        const float th = tanhf(v);
        v = (s - th) + th;  // (sign - tanh).detach() + tanh, model/model.py:137
      }
      crow[col] = v;
    }
  }
}

__device__ __forceinline__ void epi_pool(const GemmProblem& g, int row, int n0, int c, const uint32_t (&r)[32]) {
  // output rows of this input row: a query -> one row after the product occurrences; a product -> one per occurrence
  int t0, t1;
  if (g.pool_is_product) {
    t0 = g.pool_prefix[row];
    t1 = g.pool_prefix[row + 1];
  } else {
    t0 = g.pool_row0 + row;
    t1 = t0 + 1;
  }
  const int col0 = n0 + c * 32;
  const int W = g.pool_lin_w + g.pool_msl;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int col = col0 + i;
    v[i] = col < g.pool_lin_w ? tanhf(__uint_as_float(r[i]) + g.bias[col]) : 0.0f;
  }
  for (int t = t0; t < t1; ++t) {
    const int64_t pos = g.pool_pos[g.pool_is_product ? t : row];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int col = col0 + i;
      if (col >= W) continue;
      const float x = col < g.pool_lin_w ? v[i] : tanhf(g.pool_pe[pos * g.pool_msl + (col - g.pool_lin_w)]);
      g.pool_U[(size_t)t * W + col] = x;
      store_hilo(g.out_hi, g.out_lo, (size_t)t * g.out_ld + col, x);
    }
  }
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
  const int which = (int)blockIdx.x >= L.tiles0 ? 1 : 0;
  const GemmProblem& g = L.p[which];
  const CUtensorMap* tm_ah = which ? &tm_ah1 : &tm_ah0;
  const CUtensorMap* tm_al = which ? &tm_al1 : &tm_al0;
  const CUtensorMap* tm_bh = which ? &tm_bh1 : &tm_bh0;
  const CUtensorMap* tm_bl = which ? &tm_bl1 : &tm_bl0;
  const int tile = (int)blockIdx.x - (which ? L.tiles0 : 0);
  const int tile_n = tile % g.tiles_n, tile_m = tile / g.tiles_n;
  const int m0 = tile_m * kTileQ;
  const int n0 = tile_n * g.bn;
  const uint32_t b_bytes = (uint32_t)g.bn * 128u;  // one B box: bn rows x 64 bf16

  const uint32_t bar_base = smem_base + (uint32_t)(kGemmStages * kGemmStageBytes);
  const uint32_t full_bar = bar_base;             // [kGemmStages]
  const uint32_t empty_bar = bar_base + 64;       // [kGemmStages]
  const uint32_t tfull_bar = bar_base + 128;      // [1]
  const uint32_t tmem_ptr_addr = bar_base + 136;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tm_ah);
    tma_prefetch_desc(tm_al);
    tma_prefetch_desc(tm_bh);
    tma_prefetch_desc(tm_bl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(tfull_bar, 1);
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
      for (int kb = 0; kb < g.num_kb; ++kb) {
        mbar_wait(empty_bar + 8 * stage, phase ^ 1u, L.err_flag, 501);
        mbar_expect_tx(full_bar + 8 * stage, 2u * (uint32_t)kKBlockBytes + 2u * b_bytes);
        const uint32_t dst = smem_base + (uint32_t)(stage * kGemmStageBytes);
        tma_load_2d(dst, tm_ah, full_bar + 8 * stage, g.a_k0 + kb * 64, m0);
        tma_load_2d(dst + kKBlockBytes, tm_al, full_bar + 8 * stage, g.a_k0 + kb * 64, m0);
        tma_load_2d(dst + 2 * kKBlockBytes, tm_bh, full_bar + 8 * stage, kb * 64, n0);
        tma_load_2d(dst + 3 * kKBlockBytes, tm_bl, full_bar + 8 * stage, kb * 64, n0);
        if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.bn >> 3) << 17) | ((uint32_t)(kTileQ >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < g.num_kb; ++kb) {
      mbar_wait(full_bar + 8 * stage, phase, L.err_flag, 503);
      tc_fence_after();
      const uint32_t sbase = smem_base + (uint32_t)(stage * kGemmStageBytes);
      const uint64_t ah = umma_desc_sw128(sbase);
      const uint64_t al = umma_desc_sw128(sbase + kKBlockBytes);
      const uint64_t bh = umma_desc_sw128(sbase + 2 * kKBlockBytes);
      const uint64_t bl = umma_desc_sw128(sbase + 3 * kKBlockBytes);
      // the last K block may hold fewer than four 16-wide slices of real columns: what follows them in the A buffer
      // belongs to the next layer's slice and must not enter the product
      const int n4 = kb + 1 == g.num_kb ? g.last_k4 : 4;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        if (k4 < n4) {
          const uint64_t o = (uint64_t)(2 * k4);
          // small terms first, the dominant product last (all of them accumulate in fp32 anyway)
          umma_bf16_n(tmem_base, al + o, bh + o, (kb | k4) != 0 ? 1u : 0u, idesc);
          umma_bf16_n(tmem_base, ah + o, bl + o, 1u, idesc);
          umma_bf16_n(tmem_base, ah + o, bh + o, 1u, idesc);
        }
      }
      umma_commit(empty_bar + 8 * stage);
      if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
    }
    umma_commit(tfull_bar);
  } else if (warp >= 4) {
    // ===================== epilogue: a thread owns one output row =====================
    const int quarter = warp & 3;
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < g.M;
    mbar_wait(tfull_bar, 0, L.err_flag, 505);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    if (g.epi == EPI_GRU) {
      // tile = 32 hidden units x (r | z | n): columns [0, 32) r, [32, 64) z, [64, 96) n
      uint32_t rr[32], rz[32], rn[32];
      tmem_ld32(taddr, rr);
      tmem_ld32(taddr + 32u, rz);
      tmem_ld32(taddr + 64u, rn);
      tmem_ld_wait();
      if (row_ok) {
        const int H = g.gru_H;
        const int u0 = tile_n * 32;
        const float* ghr = g.gru_gh + (size_t)row * (size_t)g.gru_gh_ld;
        const float* xr = g.gru_x + (size_t)row * (size_t)g.gru_x_ld;
        const float* gpr = g.gru_gp + (size_t)row * (size_t)H;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
          const int u = u0 + i;
          if (u >= H) break;
          const float ir = __uint_as_float(rr[i]) + g.gru_b_ih[u], iz = __uint_as_float(rz[i]) + g.gru_b_ih[H + u],
                      in_ = __uint_as_float(rn[i]) + g.gru_b_ih[2 * H + u];
          const float hr = ghr[u] + g.gru_b_hh[u], hz = ghr[H + u] + g.gru_b_hh[H + u], hn = ghr[2 * H + u] + g.gru_b_hh[2 * H + u];
          const float rg = 1.0f / (1.0f + expf(-(ir + hr)));
          const float zg = 1.0f / (1.0f + expf(-(iz + hz)));
          const float ng = tanhf(in_ + rg * hn);
          const float x = u < g.gru_in_w ? xr[u] : 0.0f;
          const float h = (1.0f - zg) * ng + zg * x;
          const float o = fmaxf(gpr[u] + h, 0.0f);  // HeteroConv sum of the GAT and GatedGraphConv branches, relu
          g.C[(size_t)row * (size_t)g.ldc + u] = o;
          store_hilo(g.out_hi, g.out_lo, (size_t)row * g.out_ld + g.out_k0 + u, o);
        }
      }
    } else {
      float att_partial = 0.0f;
      const int n_chunks = g.bn / 32;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (!row_ok) continue;
        if (g.epi == EPI_STORE) {
          epi_store(g, row, n0, c, r);
        } else if (g.epi == EPI_ATT) {
          // plain fp32 store (16-byte vectors: ldc and n0 are multiples of 128) + the tile's share of <C[row, part], att>
          float* crow = g.C + (size_t)row * (size_t)g.ldc + n0 + c * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<uint4*>(crow)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          const int part = tile_n / g.att_tiles_per_part;
          const float* att = part < 4 ? g.att[part] : nullptr;
          if (att != nullptr) {
            const int pc0 = (tile_n - part * g.att_tiles_per_part) * 128 + c * 32;  // column inside the part
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (pc0 + i < g.att_width) att_partial = fmaf(__uint_as_float(r[i]), att[pc0 + i], att_partial);
          }
        } else if (g.epi == EPI_POOL) {
          epi_pool(g, row, n0, c, r);
        } else {  // EPI_ATTPOOL
          const int col0 = n0 + c * 32;
          const float* bc = g.ap_bc + (size_t)g.ap_node_graph[row] * (size_t)g.N;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = col0 + i;
            if (col < g.N)
              att_partial += g.ap_w[col] * (1.0f / (1.0f + expf(-(__uint_as_float(r[i]) + g.bias[col] + bc[col]))));
          }
        }
      }
      if (row_ok && g.epi == EPI_ATT) {
        const int part = tile_n / g.att_tiles_per_part;
        if (part < 4 && g.att[part] != nullptr)
          g.att_out[part][(size_t)row * g.att_tiles_per_part + (tile_n - part * g.att_tiles_per_part)] = att_partial;
      }
      if (row_ok && g.epi == EPI_ATTPOOL) g.ap_out[(size_t)row * g.tiles_n + tile_n] = att_partial;
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
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
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
  if (total <= 0) return 0;
  static SmemAttr attr;  // per device
  if (attr.ensure(gemm_bf16x3_kernel, kGemmSmemBytes)) return 1;
  gemm_bf16x3_kernel<<<(unsigned)total, kGemmThreads, kGemmSmemBytes, stream>>>(
      *(const CUtensorMap*)tm[0], *(const CUtensorMap*)tm[1], *(const CUtensorMap*)tm[2], *(const CUtensorMap*)tm[3],
      *(const CUtensorMap*)tm[4], *(const CUtensorMap*)tm[5], *(const CUtensorMap*)tm[6], *(const CUtensorMap*)tm[7], L);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
