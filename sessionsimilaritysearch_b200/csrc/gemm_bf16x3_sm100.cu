// gemm_bf16x3_sm100.cu — the encoder's dense linears on the tensor cores with fp32-level accuracy.
//
// C[M, N] (fp32, row stride ldc) = A[M, K] * B[N, K]^T with both operands split into two bf16 terms,
// x = hi + lo (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits together), and three tcgen05 products accumulated
// in fp32 in TMEM:  A*B ~= Ah*Bh + Ah*Bl + Al*Bh   (the dropped Al*Bl term and the split residuals are ~2^-16
// relative).  Through the whole encoder (three HeteroGGNN layers + pooling, model shape 768 -> 3 x 800 -> 3168 ->
// 1600) this stays 2.7e-5 of the output scale from a float64 forward, ten times inside the parity tolerance
// (pedantic fp32: 2.5e-6).  Replaces the x @ W.T of torch.nn.Linear / GATConv.lin_src / GRUCell inside
// model/gnn.py:64-81,193-217 that sss_encoder_forward otherwise sends to cuBLAS' SIMT sgemm (83 % of the forward).
//
// One CTA per 128 x 128 output tile, 256 threads, warp-specialised like the scan kernels:
//   warp 0  TMA producer: per 64-wide K block one stage [Ah | Al | Bh | Bl] of four 16 KB SWIZZLE_128B boxes
//   warp 1  MMA issuer: 3 products x 4 (K = 16) tcgen05.mma kind::f16 per stage into ONE 128-column accumulator
//   warp 2  TMEM allocator
//   warps 4-7  epilogue: tcgen05.ld 32 columns at a time, bounds-checked fp32 stores (a thread owns one row)
// The stage is 64 KB for 768 tensor cycles of MMA (85 B/clk/SM), so the kernel is bound by the L2 -> SM feed, not by
// the tensor pipe; that is still several times the SIMT rate, and these GEMMs are small (one wave of tiles).
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

struct GemmParams {
  float* C;
  int M, N, ldc, num_kb;
  int* err_flag;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
                   const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bl,
                   const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = (int)blockIdx.x * kTileQ;
  const int n0 = (int)blockIdx.y * kTileRows;

  const uint32_t bar_base = smem_base + (uint32_t)(kGemmStages * kGemmStageBytes);
  const uint32_t full_bar = bar_base;             // [kGemmStages]
  const uint32_t empty_bar = bar_base + 64;       // [kGemmStages]
  const uint32_t tfull_bar = bar_base + 128;      // [1]
  const uint32_t tmem_ptr_addr = bar_base + 136;
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_ah);
    tma_prefetch_desc(&tm_al);
    tma_prefetch_desc(&tm_bh);
    tma_prefetch_desc(&tm_bl);
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
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(empty_bar + 8 * stage, phase ^ 1u, p.err_flag, 501);
        mbar_expect_tx(full_bar + 8 * stage, (uint32_t)kGemmStageBytes);
        const uint32_t dst = smem_base + (uint32_t)(stage * kGemmStageBytes);
        tma_load_2d(dst, &tm_ah, full_bar + 8 * stage, kb * 64, m0);
        tma_load_2d(dst + kKBlockBytes, &tm_al, full_bar + 8 * stage, kb * 64, m0);
        tma_load_2d(dst + 2 * kKBlockBytes, &tm_bh, full_bar + 8 * stage, kb * 64, n0);
        tma_load_2d(dst + 3 * kKBlockBytes, &tm_bl, full_bar + 8 * stage, kb * 64, n0);
        if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < p.num_kb; ++kb) {
      mbar_wait(full_bar + 8 * stage, phase, p.err_flag, 503);
      tc_fence_after();
      const uint32_t sbase = smem_base + (uint32_t)(stage * kGemmStageBytes);
      const uint64_t ah = umma_desc_sw128(sbase);
      const uint64_t al = umma_desc_sw128(sbase + kKBlockBytes);
      const uint64_t bh = umma_desc_sw128(sbase + 2 * kKBlockBytes);
      const uint64_t bl = umma_desc_sw128(sbase + 3 * kKBlockBytes);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint64_t o = (uint64_t)(2 * k4);
        // small terms first, the dominant product last (all of them accumulate in fp32 anyway)
        umma_bf16(tmem_base, al + o, bh + o, (kb | k4) != 0 ? 1u : 0u);
        umma_bf16(tmem_base, ah + o, bl + o, 1u);
        umma_bf16(tmem_base, ah + o, bh + o, 1u);
      }
      umma_commit(empty_bar + 8 * stage);
      if (++stage == kGemmStages) { stage = 0; phase ^= 1u; }
    }
    umma_commit(tfull_bar);
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> global, a thread owns one output row =====================
    const int quarter = warp & 3;
    const int row = m0 + quarter * 32 + lane;
    mbar_wait(tfull_bar, 0, p.err_flag, 505);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* crow = p.C + (size_t)row * (size_t)p.ldc;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(taddr + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      if (row < p.M) {
        const int col0 = n0 + c * 32;
        if (col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(crow + col0) & 15) == 0)) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<uint4*>(crow + col0)[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i < p.N) crow[col0 + i] = __uint_as_float(r[i]);
        }
      }
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
// hi, lo bf16 [rows_pad, cols_pad] zero padded
__global__ void split_bf16_kernel(const float* __restrict__ x, int rows, int cols, int64_t ld, int transposed,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int rows_pad,
                                  int cols_pad) {
  const int64_t total = (int64_t)rows_pad * cols_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols_pad), c = (int)(i % cols_pad);
    float v = 0.0f;
    if (r < rows && c < cols) v = transposed ? x[(int64_t)c * ld + r] : x[(int64_t)r * ld + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace

int launch_split_bf16(const float* x, int rows, int cols, int64_t ld, int transposed, void* hi, void* lo, int rows_pad,
                      int cols_pad, cudaStream_t stream) {
  // (rows_pad is the number of rows WRITTEN, zero beyond `rows`: a whole operand, or one part of a fused weight)
  SSS_REQUIRE(cols_pad % 64 == 0 && rows_pad >= rows && cols_pad >= cols, "split_bf16: bad padded shape");
  const int64_t total = (int64_t)rows_pad * cols_pad;
  if (total == 0) return 0;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  split_bf16_kernel<<<blocks, 256, 0, stream>>>(x, rows, cols, ld, transposed, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                rows_pad, cols_pad);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gemm_bf16x3(const void* a_hi, const void* a_lo, int m_pad, const void* b_hi, const void* b_lo, int n_pad,
                       int k_pad, float* C, int M, int N, int ldc, int* err_flag, cudaStream_t stream) {
  SSS_REQUIRE(m_pad % 128 == 0 && n_pad % 128 == 0 && k_pad % 64 == 0 && k_pad > 0, "gemm_bf16x3: bad padded shape");
  SSS_REQUIRE(M <= m_pad && N <= n_pad, "gemm_bf16x3: output larger than the padded operands");
  if (M <= 0 || N <= 0) return 0;
  alignas(64) unsigned char tm[4][128];
  if (make_tensor_map_bf16_2d(tm[0], a_hi, (uint64_t)m_pad, (uint64_t)k_pad, 128)) return 1;
  if (make_tensor_map_bf16_2d(tm[1], a_lo, (uint64_t)m_pad, (uint64_t)k_pad, 128)) return 1;
  if (make_tensor_map_bf16_2d(tm[2], b_hi, (uint64_t)n_pad, (uint64_t)k_pad, 128)) return 1;
  if (make_tensor_map_bf16_2d(tm[3], b_lo, (uint64_t)n_pad, (uint64_t)k_pad, 128)) return 1;
  static SmemAttr attr;  // per device
  if (attr.ensure(gemm_bf16x3_kernel, kGemmSmemBytes)) return 1;
  GemmParams p;
  p.C = C;
  p.M = M;
  p.N = N;
  p.ldc = ldc;
  p.num_kb = k_pad / 64;
  p.err_flag = err_flag;
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((N + 127) / 128));
  gemm_bf16x3_kernel<<<grid, kGemmThreads, kGemmSmemBytes, stream>>>(
      *(const CUtensorMap*)tm[0], *(const CUtensorMap*)tm[1], *(const CUtensorMap*)tm[2], *(const CUtensorMap*)tm[3], p);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
