// scan_fp32.cu — the bit-faithful fp32 scan (SSS_MODE_FP32): every score is
//   acc = fmaf(q[j], x[j], acc), j = 0..d-1            (inner product)
//   acc = fmaf(q[j]-x[j], q[j]-x[j], acc)              (squared L2)
// on CUDA cores, one accumulator per (query, row) pair, k strictly ascending — the same sequence of
// roundings as oracle/search_oracle.c.  Replaces faiss IndexFlatIP/IndexFlatL2.search
// (test_amazon_filterd.py:211-220,578) with a defined summation order.
//
// Tiling: 64 queries x 64 rows per 256-thread block, 4x4 accumulators per thread, K staged through
// shared memory 32 columns at a time (zero padded: fmaf(0,0,acc) == acc, so padding is exact).
#include "common.cuh"
#include "kernels.h"

namespace sss {

constexpr int FT = 64;   // tile edge
constexpr int FK = 32;   // k chunk

template <int METRIC>
__global__ void __launch_bounds__(256) scan_fp32_kernel(const float* __restrict__ db, int d, int64_t row_begin,
                                                        int64_t row_end, const float* __restrict__ q, int64_t nq,
                                                        SelectState st) {
  __shared__ __align__(16) float Qs[FK][FT + 4];
  __shared__ __align__(16) float Ds[FK][FT + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t row0 = row_begin + (int64_t)blockIdx.x * FT;
  const int64_t q0 = (int64_t)blockIdx.y * FT;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < d; k0 += FK) {
    // 64 rows x 32 k of each operand, transposed into [k][row]
    for (int i = tid; i < FT * FK; i += 256) {
      int r = i / FK, kk = i % FK;
      int64_t gq = q0 + r, gr = row0 + r;
      int gk = k0 + kk;
      Qs[kk][r] = (gq < nq && gk < d) ? q[gq * (int64_t)d + gk] : 0.0f;
      Ds[kk][r] = (gr < row_end && gk < d) ? db[gr * (int64_t)d + gk] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < FK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&Qs[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Ds[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (METRIC == 0) {
            acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
          } else {
            float t = __fsub_rn(a[i], b[j]);
            acc[i][j] = __fmaf_rn(t, t, acc[i][j]);
          }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gq = q0 + ty * 4 + i;
    if (gq >= nq) continue;
    const float thr = st.thr[gq];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gr = row0 + tx * 4 + j;
      if (gr >= row_end) continue;
      const float s = METRIC == 0 ? acc[i][j] : -acc[i][j];
      if (s > thr) {
        uint32_t slot = atomicAdd(&st.cnt[gq], 1u);
        if (slot < (uint32_t)st.cap) st.cand[(size_t)gq * st.cap + slot] = pack_cand(score_key(s), (uint32_t)gr);
      }
    }
  }
}

int launch_scan_fp32(const float* db, int d, int metric, int64_t row_begin, int64_t row_end, const float* q, int64_t nq,
                     SelectState st, cudaStream_t stream) {
  if (row_end <= row_begin || nq <= 0) return 0;
  dim3 grid((unsigned)((row_end - row_begin + FT - 1) / FT), (unsigned)((nq + FT - 1) / FT));
  SSS_REQUIRE(grid.y <= 65535, "too many queries in one search call (max 4M)");
  if (metric == 0)
    scan_fp32_kernel<0><<<grid, 256, 0, stream>>>(db, d, row_begin, row_end, q, nq, st);
  else
    scan_fp32_kernel<1><<<grid, 256, 0, stream>>>(db, d, row_begin, row_end, q, nq, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
