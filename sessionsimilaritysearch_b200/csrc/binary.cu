// binary.cu — binary-hash retrieval: sign-binarise + pack, and the Hamming scan.
// Replaces (d_emb+1)/2 -> astype(int) -> np.packbits -> faiss.IndexBinaryFlat.search
// (fine_tune_ours.py:839-843,871-876).  Scores are -distance as exact small integers in fp32, so the
// shared filter/refine machinery (select.cu) applies unchanged; ties (the norm for integer distances) go
// to the smaller id.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace sss {

// bit = (x > 0), MSB first within each byte, zero padded (np.packbits semantics on {0,1} input)
__global__ void pack_sign_bits_kernel(const float* __restrict__ x, uint8_t* __restrict__ codes, int64_t n, int nbits,
                                      int nbytes) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * nbytes) return;
  int64_t row = t / nbytes;
  int b = (int)(t % nbytes);
  unsigned v = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int col = b * 8 + j;
    unsigned bit = (col < nbits && x[row * (int64_t)nbits + col] > 0.0f) ? 1u : 0u;
    v = (v << 1) | bit;
  }
  codes[t] = (uint8_t)v;
}

int launch_pack_sign_bits(const float* x, uint8_t* codes, int64_t n, int nbits, cudaStream_t st) {
  int nbytes = (nbits + 7) / 8;
  int64_t total = n * nbytes;
  if (total <= 0) return 0;
  pack_sign_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, codes, n, nbits, nbytes);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- tensor-core form: codes as +-1.0 in E4M3 ------------------------------------------------------------------
// Every code bit becomes one fp8 byte (1 -> +1.0 = 0x38, 0 -> -1.0 = 0xB8), rows padded with 0x00 (= 0.0) to a
// multiple of 128 bytes.  Then <a, b> = nbits - 2 * hamming(a, b), every product is +-1 and the fp32 accumulation of
// <= 256 of them is exact: the fused scan + top-k machinery of the float index applies unchanged with integer scores
// (ties, the norm here, go to the smaller id).  8x the bytes of the packed codes, for the tensor cores' rate: a
// popcount scan manages ~5e11 (query, row) pairs/s on this chip (16 POPC per clock per SM, 8 per pair), the fp8 MMA 5e12.
__global__ void expand_codes_fp8_kernel(const uint8_t* __restrict__ codes, int nbytes, int64_t n, int row_bytes,
                                        uint8_t* __restrict__ out) {
  // one thread per output group of 8 bytes (= one code byte)
  const int groups = row_bytes / 8;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * groups) return;
  const int64_t row = t / groups;
  const int g = (int)(t % groups);
  uint2 v = make_uint2(0u, 0u);
  if (g < nbytes) {
    const uint32_t c = codes[row * nbytes + g];  // bit 7 is the first code bit (np.packbits is MSB first)
    auto four = [](uint32_t b4) {  // 4 bits (MSB first) -> 4 bytes in memory order
      uint32_t w = 0xB8B8B8B8u;
      // +1.0 (0x38) differs from -1.0 (0xB8) in the sign bit only: clear it where the bit is set
      w ^= ((b4 >> 3) & 1u) * 0x00000080u;
      w ^= ((b4 >> 2) & 1u) * 0x00008000u;
      w ^= ((b4 >> 1) & 1u) * 0x00800000u;
      w ^= (b4 & 1u) * 0x80000000u;
      return w;
    };
    v.x = four(c >> 4);
    v.y = four(c & 15u);
  }
  reinterpret_cast<uint2*>(out)[t] = v;
}

int launch_expand_codes_fp8(const uint8_t* codes, int nbytes, int64_t n, int row_bytes, uint8_t* out, cudaStream_t st) {
  SSS_REQUIRE(row_bytes % 128 == 0 && row_bytes >= nbytes * 8, "expand_codes_fp8: row pitch must hold 8 bytes per code byte");
  const int64_t total = n * (row_bytes / 8);
  if (total <= 0) return 0;
  expand_codes_fp8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(codes, nbytes, n, row_bytes, out);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// query staging of a binary search: fp8 expansion of the query codes (zero rows for padding queries), packed copy for
// the popcount scan, selection state reset.  One block per query row.
__global__ void prep_binary_kernel(const uint8_t* __restrict__ q_codes, int64_t nq, int nbytes, int row_bytes,
                                   uint8_t* __restrict__ q_fp8, uint8_t* __restrict__ q_packed, int pitch,
                                   SelectState st) {
  const int64_t row = blockIdx.x;
  for (int g = threadIdx.x; g < row_bytes / 8; g += blockDim.x) {
    uint2 v = make_uint2(0u, 0u);
    if (row < nq && g < nbytes) {
      const uint32_t c = q_codes[row * nbytes + g];
      uint32_t lo = 0xB8B8B8B8u, hi = 0xB8B8B8B8u;
      lo ^= ((c >> 7) & 1u) * 0x00000080u; lo ^= ((c >> 6) & 1u) * 0x00008000u;
      lo ^= ((c >> 5) & 1u) * 0x00800000u; lo ^= ((c >> 4) & 1u) * 0x80000000u;
      hi ^= ((c >> 3) & 1u) * 0x00000080u; hi ^= ((c >> 2) & 1u) * 0x00008000u;
      hi ^= ((c >> 1) & 1u) * 0x00800000u; hi ^= (c & 1u) * 0x80000000u;
      v = make_uint2(lo, hi);
    }
    if (q_fp8) reinterpret_cast<uint2*>(q_fp8 + row * (int64_t)row_bytes)[g] = v;
  }
  if (q_packed && row < nq)
    for (int b = threadIdx.x; b < pitch; b += blockDim.x) q_packed[row * pitch + b] = b < nbytes ? q_codes[row * nbytes + b] : (uint8_t)0;
  if (threadIdx.x == 0) {
    st.margin[row] = 0.0f;
    if (st.qn2 != nullptr) st.qn2[row] = 0.0f;
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

const void* prep_binary_kernel_addr() { return (const void*)prep_binary_kernel; }

int launch_prep_binary(const uint8_t* q_codes, int64_t nq, int64_t nq_pad, int nbytes, int row_bytes, uint8_t* q_fp8,
                       uint8_t* q_packed, int pitch, SelectState st, cudaStream_t stream) {
  prep_binary_kernel<<<(unsigned)nq_pad, 64, 0, stream>>>(q_codes, nq, nbytes, row_bytes, q_fp8, q_packed, pitch, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// One thread per DB row (code held in registers as 32-bit words), a tile of 64 queries in shared memory.
// Codes are stored padded to a multiple of 4 bytes (nwords words per row).
constexpr int HQ = 64;
constexpr int HMAXW = 16;  // up to 512 bits
__global__ void __launch_bounds__(256) scan_hamming_kernel(const uint32_t* __restrict__ db, int nwords,
                                                           int64_t row_begin, int64_t row_end,
                                                           const uint32_t* __restrict__ q, int64_t nq, SelectState st) {
  __shared__ uint32_t qs[HQ][HMAXW];
  __shared__ float thr_s[HQ];
  const int64_t q0 = (int64_t)blockIdx.y * HQ;
  for (int i = threadIdx.x; i < HQ * nwords; i += blockDim.x) {
    int qi = i / nwords, w = i % nwords;
    qs[qi][w] = (q0 + qi < nq) ? q[(q0 + qi) * nwords + w] : 0u;
  }
  for (int i = threadIdx.x; i < HQ; i += blockDim.x) thr_s[i] = (q0 + i < nq) ? st.thr[q0 + i] : INFINITY;
  __syncthreads();
  const int64_t row = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= row_end) return;
  uint32_t x[HMAXW];
#pragma unroll
  for (int w = 0; w < HMAXW; ++w) x[w] = w < nwords ? db[row * nwords + w] : 0u;
  for (int qi = 0; qi < HQ; ++qi) {
    int dist = 0;
#pragma unroll
    for (int w = 0; w < HMAXW; ++w)
      if (w < nwords) dist += __popc(x[w] ^ qs[qi][w]);
    const float s = -(float)dist;
    if (s > thr_s[qi]) {
      const int64_t gq = q0 + qi;
      uint32_t slot = atomicAdd(&st.cnt[gq], 1u);
      if (slot < (uint32_t)st.cap) st.cand[(size_t)gq * st.cap + slot] = pack_cand(score_key(s), (uint32_t)row);
    }
  }
}

// Few queries (<= kHammingSmallNq): the packed codes are the cheapest thing to stream (32 bytes per 256-bit code, an
// eighth of the +-1 E4M3 rows of the tensor path), and popcounts against a handful of queries cost less than the
// stream.  Grid-stride, a thread per row, 16-byte loads, only the real queries in the inner loop.
template <int NW>  // words per code, a multiple of 4
__global__ void __launch_bounds__(256) scan_hamming_small_kernel(const uint32_t* __restrict__ db, int64_t row_begin,
                                                                 int64_t row_end, const uint32_t* __restrict__ q,
                                                                 int nq, SelectState st) {
  __shared__ uint32_t qs[kHammingSmallNq][NW];
  __shared__ float thr_s[kHammingSmallNq];
  for (int i = threadIdx.x; i < nq * NW; i += blockDim.x) qs[i / NW][i % NW] = q[i];
  for (int i = threadIdx.x; i < nq; i += blockDim.x) thr_s[i] = st.thr[i];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < row_end; row += stride) {
    uint32_t x[NW];
    const uint4* p = reinterpret_cast<const uint4*>(db + row * NW);
#pragma unroll
    for (int v = 0; v < NW / 4; ++v) {
      const uint4 u = __ldg(p + v);
      x[4 * v] = u.x; x[4 * v + 1] = u.y; x[4 * v + 2] = u.z; x[4 * v + 3] = u.w;
    }
    for (int qi = 0; qi < nq; ++qi) {
      int dist = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) dist += __popc(x[w] ^ qs[qi][w]);
      const float s = -(float)dist;
      if (s > thr_s[qi]) {
        const uint32_t slot = atomicAdd(&st.cnt[qi], 1u);
        if (slot < (uint32_t)st.cap) st.cand[(size_t)qi * st.cap + slot] = pack_cand(score_key(s), (uint32_t)row);
      }
    }
  }
}

template <int NW>
__device__ __forceinline__ int hamming_words(const uint32_t (&x)[NW], const uint32_t (&y)[NW]) {
  // (a carry-save adder tree — 4 POPC + 22 logic operations per 256 bits instead of 8 POPC — was measured: no gain,
  // logic instructions run at half rate and the tree needs as many cycles as the quarter-rate POPCs it saves)
  int dist = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) dist += __popc(x[w] ^ y[w]);
  return dist;
}

// Up to 8 queries with NW * NQ <= 64: the query words and thresholds live in REGISTERS (no shared-memory reads in the
// pair loop) and a thread works on two rows at a time (their loads overlap).
template <int NW, int NQ>
__global__ void __launch_bounds__(256) scan_hamming_reg_kernel(const uint32_t* __restrict__ db, int64_t row_begin,
                                                               int64_t row_end, const uint32_t* __restrict__ q, int nq,
                                                               SelectState st) {
  uint32_t qr[NQ][NW];
  float thr[NQ];
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
    for (int w = 0; w < NW; ++w) qr[qi][w] = qi < nq ? __ldg(q + qi * NW + w) : 0u;
    thr[qi] = qi < nq ? st.thr[qi] : INFINITY;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = row_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < row_end; row += 2 * stride) {
    const int64_t row_b = row + stride;
    const bool have_b = row_b < row_end;
    uint32_t xa[NW], xb[NW];
    const uint4* pa = reinterpret_cast<const uint4*>(db + row * NW);
    const uint4* pb = reinterpret_cast<const uint4*>(db + (have_b ? row_b : row) * NW);
#pragma unroll
    for (int v = 0; v < NW / 4; ++v) {
      const uint4 u = __ldg(pa + v), t = __ldg(pb + v);
      xa[4 * v] = u.x; xa[4 * v + 1] = u.y; xa[4 * v + 2] = u.z; xa[4 * v + 3] = u.w;
      xb[4 * v] = t.x; xb[4 * v + 1] = t.y; xb[4 * v + 2] = t.z; xb[4 * v + 3] = t.w;
    }
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
      const float sa = -(float)hamming_words<NW>(xa, qr[qi]);
      const float sb = -(float)hamming_words<NW>(xb, qr[qi]);
      if (sa > thr[qi]) {
        const uint32_t slot = atomicAdd(&st.cnt[qi], 1u);
        if (slot < (uint32_t)st.cap) st.cand[(size_t)qi * st.cap + slot] = pack_cand(score_key(sa), (uint32_t)row);
      }
      if (have_b && sb > thr[qi]) {
        const uint32_t slot = atomicAdd(&st.cnt[qi], 1u);
        if (slot < (uint32_t)st.cap) st.cand[(size_t)qi * st.cap + slot] = pack_cand(score_key(sb), (uint32_t)row_b);
      }
    }
  }
}

template <int NW>
static bool launch_hamming_reg(unsigned grid, const uint32_t* db, int64_t row_begin, int64_t row_end, const uint32_t* q,
                               int nq, SelectState st, cudaStream_t stream) {
  if (nq == 1) scan_hamming_reg_kernel<NW, 1><<<grid, 256, 0, stream>>>(db, row_begin, row_end, q, nq, st);
  else if (nq == 2) scan_hamming_reg_kernel<NW, 2><<<grid, 256, 0, stream>>>(db, row_begin, row_end, q, nq, st);
  else if (nq <= 4) scan_hamming_reg_kernel<NW, 4><<<grid, 256, 0, stream>>>(db, row_begin, row_end, q, nq, st);
  else if (nq <= 8 && NW <= 8) scan_hamming_reg_kernel<NW, (NW <= 8 ? 8 : 4)><<<grid, 256, 0, stream>>>(db, row_begin, row_end, q, nq, st);
  else return false;
  return true;
}

int launch_scan_hamming(const uint8_t* db, int nbytes, int64_t row_begin, int64_t row_end, const uint8_t* q, int64_t nq,
                        SelectState st, cudaStream_t stream) {
  if (row_end <= row_begin || nq <= 0) return 0;
  SSS_REQUIRE(nbytes % 4 == 0 && nbytes / 4 <= HMAXW, "Hamming scan needs codes padded to 4-byte words, <= 512 bits");
  const int nw = nbytes / 4;
  if (nq <= kHammingSmallNq && (nw == 4 || nw == 8 || nw == 16)) {
    const int64_t blocks = (row_end - row_begin + 255) / 256;
    const unsigned grid = (unsigned)std::min<int64_t>(blocks, (int64_t)current_sm_count() * 8);
    const uint32_t* d32 = (const uint32_t*)db;
    const uint32_t* q32 = (const uint32_t*)q;
    const bool in_regs = nw == 4 ? launch_hamming_reg<4>(grid, d32, row_begin, row_end, q32, (int)nq, st, stream)
                       : nw == 8 ? launch_hamming_reg<8>(grid, d32, row_begin, row_end, q32, (int)nq, st, stream)
                                 : launch_hamming_reg<16>(grid, d32, row_begin, row_end, q32, (int)nq, st, stream);
    if (in_regs) {
      SSS_CUDA_OK(cudaGetLastError());
      return 0;
    }
    if (nw == 4) scan_hamming_small_kernel<4><<<grid, 256, 0, stream>>>(d32, row_begin, row_end, q32, (int)nq, st);
    else if (nw == 8) scan_hamming_small_kernel<8><<<grid, 256, 0, stream>>>(d32, row_begin, row_end, q32, (int)nq, st);
    else scan_hamming_small_kernel<16><<<grid, 256, 0, stream>>>(d32, row_begin, row_end, q32, (int)nq, st);
    SSS_CUDA_OK(cudaGetLastError());
    return 0;
  }
  dim3 grid((unsigned)((row_end - row_begin + 255) / 256), (unsigned)((nq + HQ - 1) / HQ));
  scan_hamming_kernel<<<grid, 256, 0, stream>>>((const uint32_t*)db, nbytes / 4, row_begin, row_end, (const uint32_t*)q,
                                                nq, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

// dot_bits > 0: the scores are tensor-core dot products over dot_bits +-1 elements (hamming = (dot_bits - dot) / 2);
// dot_bits == 0: the scores are -hamming from the popcount scan
__global__ void emit_hamming_kernel(SelectState st, int64_t nq, int k, int dot_bits, int64_t id_offset,
                                    int32_t* __restrict__ D, int64_t* __restrict__ I) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * k) return;
  int64_t q = t / k;
  int j = (int)(t % k);
  if ((uint32_t)j < st.nret[q]) {
    uint64_t c = st.cand[(size_t)q * st.cap + j];
    const float sc = key_score(cand_key(c));
    D[t] = dot_bits > 0 ? (int32_t)(((float)dot_bits - sc) * 0.5f) : (int32_t)(-sc);
    I[t] = (int64_t)cand_id(c) + id_offset;
  } else {
    D[t] = 0x7FFFFFFF;
    I[t] = -1;
  }
}

const void* emit_hamming_kernel_addr() { return (const void*)emit_hamming_kernel; }

int launch_emit_hamming(SelectState st, int64_t nq, int k, int dot_bits, int64_t id_offset, int32_t* D, int64_t* I,
                        cudaStream_t stream) {
  int64_t total = nq * k;
  if (total <= 0) return 0;
  emit_hamming_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(st, nq, k, dot_bits, id_offset, D, I);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
