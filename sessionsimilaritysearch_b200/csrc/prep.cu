// prep.cu — HBM-bound staging kernels: row normalisation + bf16 packing (index.add), query staging,
// segment maps and segment sums.  Replaces normalize() (util_amazon_filtered.py:28-31,
// fine_tune_ours.py:38-40) and the numpy side of build_index (test_amazon_filterd.py:207-223).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace sss {

// Rows are staged through shared memory so that global reads and writes are fully coalesced while the
// sum of squares is still accumulated by ONE thread per row in k-ascending order (the oracle's order).
__global__ void __launch_bounds__(128) add_rows_kernel(const float* __restrict__ in, int64_t n, int d, int d_pad,
                                                       int norm_mode, int rows_per_block, float* __restrict__ out_f32,
                                                       __nv_bfloat16* __restrict__ out_bf16, int64_t out_row0,
                                                       unsigned int* __restrict__ maxnorm2_bits, int aug) {
  // aug (L2 metric on the tensor path): bf16 columns d, d+1 of a row hold -||y||^2 / 2 (hi + lo) and the query side 1, 1, so
  // that the tensor-core score is <q, y> - ||y||^2 / 2 = (||q||^2 - ||q - y||^2) / 2 — ordered like -distance.
  extern __shared__ float smem[];
  const int ld = d + 1;
  float* den = smem + (size_t)rows_per_block * ld;
  float* caug = den + rows_per_block;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
  const int nrows = (int)min((int64_t)rows_per_block, n - row0);
  const int tid = threadIdx.x;
  const int total = nrows * d;
  const float* src = in + row0 * (int64_t)d;
  for (int i = tid; i < total; i += blockDim.x) smem[(i / d) * ld + (i % d)] = src[i];
  __syncthreads();
  if (tid < nrows) {
    const float* x = smem + tid * ld;
    float ss = 0.0f;
    for (int j = 0; j < d; ++j) ss = __fmaf_rn(x[j], x[j], ss);
    float dn = 1.0f, n2 = ss;
    if (norm_mode == 1) {
      dn = __fsqrt_rn(ss < 1e-6f ? 1e-6f : ss);
    } else if (norm_mode == 2) {
      dn = __fadd_rn(__fsqrt_rn(ss), 1e-4f);
    } else if (norm_mode == 3) {
      float nn = __fsqrt_rn(ss);
      dn = nn < 1e-12f ? 1e-12f : nn;
    }
    if (norm_mode != 0) n2 = ss / (dn * dn);
    den[tid] = dn;
    if (maxnorm2_bits) {
      // ||x - bf16(x)||^2 of the stored row: the exact rounding error the tensor-core scan will see
      float e2 = 0.0f, n2y = 0.0f, n4y = 0.0f;
      for (int j = 0; j < d; ++j) {
        const float y = norm_mode == 0 ? x[j] : __fdiv_rn(x[j], dn);
        const float r = y - __bfloat162float(__float2bfloat16_rn(y));
        e2 = __fmaf_rn(r, r, e2);
        n2y = __fmaf_rn(y, y, n2y);
        n4y = __fmaf_rn(y * y, y * y, n4y);
      }
      if (aug) {
        // -||y||^2 / 2 in TWO bf16 columns (hi + lo: 16 mantissa bits — one column would put 2^-9 * ||y||^2 / 2 of
        // rounding error into every score, far more than the rest of the slack for unnormalised rows)
        const float c = -0.5f * n2y;
        caug[tid] = c;
        const float chi = __bfloat162float(__float2bfloat16_rn(c));
        const float clo = __bfloat162float(__float2bfloat16_rn(c - chi));
        // what is left: the split residual + the fp32 rounding of ||y||^2 itself
        const float ec = fabsf(c - chi - clo) + (float)(d + 8) * 1.2e-7f * fabsf(c);
        e2 = __fmaf_rn(ec, ec, e2);
        n2 = n2 + chi * chi + clo * clo;
      }
      atomicMax(maxnorm2_bits, __float_as_uint(n2 * 1.00002f + 1e-30f));
      atomicMax(maxnorm2_bits + 1, __float_as_uint(e2 * 1.00002f + 1e-30f));
      atomicMax(maxnorm2_bits + 2, __float_as_uint(n4y * 1.00002f + 1e-30f));  // max ||y||_4^4 (statistical slack)
    }
  }
  __syncthreads();
  if (out_f32) {
    float* dst = out_f32 + (out_row0 + row0) * (int64_t)d;
    for (int i = tid; i < total; i += blockDim.x) {
      int r = i / d, j = i % d;
      float v = smem[r * ld + j];
      dst[i] = norm_mode == 0 ? v : __fdiv_rn(v, den[r]);
    }
  }
  if (out_bf16) {
    __nv_bfloat16* dstb = out_bf16 + (out_row0 + row0) * (int64_t)d_pad;
    const int totalp = nrows * d_pad;
    for (int i = tid; i < totalp; i += blockDim.x) {
      int r = i / d_pad, j = i % d_pad;
      float v = 0.0f;
      if (j < d) {
        v = smem[r * ld + j];
        if (norm_mode != 0) v = __fdiv_rn(v, den[r]);
      } else if (aug && j == d) {
        v = caug[r];
      } else if (aug && j == d + 1) {
        v = caug[r] - __bfloat162float(__float2bfloat16_rn(caug[r]));
      }
      dstb[i] = __float2bfloat16_rn(v);
    }
  }
}

int launch_add_rows(const float* in, int64_t n, int d, int d_pad, int norm_mode, float* out_f32, void* out_bf16,
                    int64_t out_row0, unsigned int* maxnorm2_bits, cudaStream_t st, int aug) {
  SSS_REQUIRE(!aug || (out_bf16 != nullptr && maxnorm2_bits != nullptr && d_pad >= d + 2),
              "add_rows: the L2 columns need the bf16 copy, its statistics and two spare columns");
  if (n <= 0) return 0;
  int rpb = (int)(60 * 1024 / (sizeof(float) * (d + 1)));
  if (rpb > 128) rpb = 128;
  if (rpb < 1) rpb = 1;
  size_t smem = sizeof(float) * ((size_t)rpb * (d + 1) + 2 * (size_t)rpb);
  SSS_REQUIRE(smem <= 200 * 1024, "embedding width too large for add_rows_kernel");
  static SmemAttr attr;
  if (attr.ensure(add_rows_kernel, (int)smem)) return 1;
  int64_t blocks = (n + rpb - 1) / rpb;
  add_rows_kernel<<<(unsigned)blocks, 128, smem, st>>>(in, n, d, d_pad, norm_mode, rpb, out_f32,
                                                        (__nv_bfloat16*)out_bf16, out_row0, maxnorm2_bits, aug);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// Query staging: bf16 copy (zero padded to [nq_pad, d_pad]), private fp32 copy, filter slack, selection state reset.
// One warp per query row.
//
// Filter slack.  Let s be the fixed-order fp32 score of (q, x) and b the tensor-core score of the bf16
// copies (q^, x^).  q.x - q^.x^ = q.(x - x^) + (q - q^).x^.
//   slack == 1 (SSS_MODE_EXACT): the rigorous bound
//       |s - b| <= ||q|| * max_rows ||x - x^||  +  ||q - q^|| * max_rows ||x^||  +  accumulation slack,
//     with the two maxima measured exactly when rows are added (stats[1], stats[0]) and the query terms measured
//     here; the accumulation slack covers d fp32 roundings on our side and a generous 8x that inside the tensor
//     core.  About 2.3x tighter than 2^-7 * ||q|| * ||x||.
//   slack == 2 (SSS_MODE_BF16): a statistical bound.  The bf16 rounding error of element i is at most 2^-9 |x_i|;
//     modelled as independent and uniform in that range, s - b has variance
//         <= (2^-18 / 3) * (sum q_i^2 x_i^2 + sum q_i^2 x^_i^2) <= (2^-17 / 3) * ||q||_4^2 * max_rows ||x||_4^2
//     (Cauchy-Schwarz; max ||x||_4^4 measured at add time, stats[2]).  The slack is 4 of those standard deviations
//     plus the accumulation term, never more than the rigorous bound.  For embedding-like rows it is ~2x tighter at
//     d = 128 and ~10x at d = 1600 (the rigorous bound grows like sqrt(d) relative to the actual noise).
__global__ void prep_queries_kernel(const float* __restrict__ q, int64_t nq, int64_t nq_pad, int d, int d_pad,
                                    __nv_bfloat16* __restrict__ q_bf16, int slack, const unsigned int* stats,
                                    SelectState st, int aug, float* __restrict__ q_keep) {
  const int warps_per_block = blockDim.x / 32;
  const int64_t row = (int64_t)blockIdx.x * warps_per_block + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= nq_pad) return;
  float ss = 0.0f, ee = 0.0f, s4 = 0.0f;
  for (int j = lane; j < d_pad; j += 32) {
    float v = (row < nq && j < d) ? q[row * (int64_t)d + j] : 0.0f;
    if (q_keep != nullptr && j < d) q_keep[row * (int64_t)d + j] = v;
    const bool one = aug && (j == d || j == d + 1) && row < nq;  // L2 on the tensor path: query side of the extra columns
    if (one) v = 1.0f;
    const __nv_bfloat16 vb = __float2bfloat16_rn(v);
    const float r = v - __bfloat162float(vb);
    if (!one) {
      ss += v * v;
      s4 += (v * v) * (v * v);
    }
    ee += r * r;
    if (q_bf16) q_bf16[row * (int64_t)d_pad + j] = vb;
  }
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    ee += __shfl_xor_sync(0xffffffffu, ee, o);
    s4 += __shfl_xor_sync(0xffffffffu, s4, o);
  }
  if (lane == 0) {
    float m = 0.0f;
    if (slack != 0 && stats != nullptr) {
      const float xn = sqrtf(__uint_as_float(stats[0]));   // max ||x||   (inflated at add time)
      const float xe = sqrtf(__uint_as_float(stats[1]));   // max ||x - x^||
      const float qn = sqrtf((ss + (aug ? 2.0f : 0.0f)) * 1.0001f), qe = sqrtf(ee * 1.0001f);
      const float acc_slack = (float)d * 5.4e-7f * qn * xn + 1e-30f;
      m = (qn * xe + qe * (xn + xe)) * 1.0001f + acc_slack;
      // (the conversion -dist = 2 * score - ||q||^2 uses this kernel's own ||q||^2: its rounding is part of the slack)
      if (aug) m += 4e-7f * (ss + 2.0f);
      if (slack == 2 && !aug) {
        const float x4 = sqrtf(sqrtf(__uint_as_float(stats[2])));  // max ||x||_4
        const float q4 = sqrtf(sqrtf(s4 * 1.0001f));
        const float sigma = 1.5946e-3f * q4 * x4;                  // sqrt(2^-17 / 3) = 1.5946e-3
        m = fminf(m, 4.0f * sigma + acc_slack);
      }
    }
    st.margin[row] = m;
    if (st.qn2 != nullptr) st.qn2[row] = ss;
    st.thr[row] = row < nq ? -INFINITY : INFINITY;
    st.cnt[row] = 0;
    st.nret[row] = 0;
    if (row == 0) {
      *st.overflow = 0;
      st.skip_cnt[0] = 0;
      st.skip_cnt[1] = 0;
    }
  }
}

const void* prep_queries_kernel_addr() { return (const void*)prep_queries_kernel; }

int launch_prep_queries(const float* q, int64_t nq, int64_t nq_pad, int d, int d_pad, void* q_bf16, int slack,
                        const unsigned int* stats, SelectState st, cudaStream_t stream, int aug, float* q_keep) {
  const int wpb = 8;
  int64_t blocks = (nq_pad + wpb - 1) / wpb;
  prep_queries_kernel<<<(unsigned)blocks, wpb * 32, 0, stream>>>(q, nq, nq_pad, d, d_pad, (__nv_bfloat16*)q_bf16,
                                                                  slack, stats, st, aug, q_keep);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// row -> segment map (segments are short: one thread per segment)
__global__ void row_seg_kernel(const int64_t* __restrict__ seg_off, int64_t n_seg, int32_t* __restrict__ row_seg) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  for (int64_t r = seg_off[s]; r < seg_off[s + 1]; ++r) row_seg[r] = (int32_t)s;
}
int launch_row_seg(const int64_t* seg_off, int64_t n_seg, int32_t* row_seg, cudaStream_t st) {
  if (n_seg <= 0) return 0;
  row_seg_kernel<<<(unsigned)((n_seg + 255) / 256), 256, 0, st>>>(seg_off, n_seg, row_seg);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// per-segment sum of rows, rows added in row order (SUM reduction is linear: <q, sum x_r> )
__global__ void segment_sum_kernel(const float* __restrict__ rows, const int64_t* __restrict__ seg_off, int64_t n_seg,
                                   int d, float* __restrict__ out) {
  int64_t s = blockIdx.x;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float acc = 0.0f;
    for (int64_t r = seg_off[s]; r < seg_off[s + 1]; ++r) acc = __fadd_rn(acc, rows[r * (int64_t)d + j]);
    out[s * (int64_t)d + j] = acc;
  }
}
int launch_segment_sum(const float* rows, const int64_t* seg_off, int64_t n_seg, int d, float* out, cudaStream_t st) {
  if (n_seg <= 0) return 0;
  segment_sum_kernel<<<(unsigned)n_seg, 128, 0, st>>>(rows, seg_off, n_seg, d, out);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// stand-alone normalize() (device in/out, may alias)
int launch_normalize(const float* in, float* out, int64_t n, int d, int norm_mode, cudaStream_t st) {
  return launch_add_rows(in, n, d, d, norm_mode, out, nullptr, 0, nullptr, st);
}

// ---- row gather: out[i, :] = table[ids[i], :] ------------------------------------------------------------
// nn.Embedding lookup (NodeAsinEmbedding.forward, model/NodeEmbedding.py:137-138) and the feature-cache gather of the
// batched featuriser.  One warp per output row, 16-byte accesses when the rows allow it; an id outside the table
// sets *bad (no silent clamping).
__global__ void gather_rows_kernel(const float* __restrict__ table, int64_t n_rows, int d, const int64_t* __restrict__ ids,
                                   int64_t n, float* __restrict__ out, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (i >= n) return;
  const int lane = threadIdx.x % 32;
  const int64_t id = ids[i];
  if (id < 0 || id >= n_rows) {
    if (lane == 0) *bad = 1;
    return;
  }
  const float* src = table + id * (int64_t)d;
  float* dst = out + i * (int64_t)d;
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int j = lane; j < d / 4; j += 32) d4[j] = __ldg(s4 + j);
  } else {
    for (int j = lane; j < d; j += 32) dst[j] = __ldg(src + j);
  }
}

int launch_gather_rows(const float* table, int64_t n_rows, int d, const int64_t* ids, int64_t n, float* out, int* bad,
                       cudaStream_t st) {
  if (n <= 0) return 0;
  const int wpb = 8;
  gather_rows_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, st>>>(table, n_rows, d, ids, n, out, bad);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
