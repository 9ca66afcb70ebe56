// binary.cu — binary-hash retrieval: sign-binarise + pack, and the Hamming scan.
// Replaces (d_emb+1)/2 -> astype(int) -> np.packbits -> faiss.IndexBinaryFlat.search
// (fine_tune_ours.py:839-843,871-876).  Scores are -distance as exact small integers in fp32, so the
// shared filter/refine machinery (select.cu) applies unchanged; ties (the norm for integer distances) go
// to the smaller id.
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

int launch_scan_hamming(const uint8_t* db, int nbytes, int64_t row_begin, int64_t row_end, const uint8_t* q, int64_t nq,
                        SelectState st, cudaStream_t stream) {
  if (row_end <= row_begin || nq <= 0) return 0;
  SSS_REQUIRE(nbytes % 4 == 0 && nbytes / 4 <= HMAXW, "Hamming scan needs codes padded to 4-byte words, <= 512 bits");
  dim3 grid((unsigned)((row_end - row_begin + 255) / 256), (unsigned)((nq + HQ - 1) / HQ));
  scan_hamming_kernel<<<grid, 256, 0, stream>>>((const uint32_t*)db, nbytes / 4, row_begin, row_end, (const uint32_t*)q,
                                                nq, st);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void emit_hamming_kernel(SelectState st, int64_t nq, int k, int64_t id_offset, int32_t* __restrict__ D,
                                    int64_t* __restrict__ I) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * k) return;
  int64_t q = t / k;
  int j = (int)(t % k);
  if ((uint32_t)j < st.nret[q]) {
    uint64_t c = st.cand[(size_t)q * st.cap + j];
    D[t] = (int32_t)(-key_score(cand_key(c)));
    I[t] = (int64_t)cand_id(c) + id_offset;
  } else {
    D[t] = 0x7FFFFFFF;
    I[t] = -1;
  }
}

int launch_emit_hamming(SelectState st, int64_t nq, int k, int nbits, int64_t id_offset, int32_t* D, int64_t* I,
                        cudaStream_t stream) {
  (void)nbits;
  int64_t total = nq * k;
  if (total <= 0) return 0;
  emit_hamming_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(st, nq, k, id_offset, D, I);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace sss
