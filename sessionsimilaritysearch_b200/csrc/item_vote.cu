// item_vote.cu — replaces get_prediction_by_knn (test_amazon_filterd.py:59-78): every item of every neighbour
// session receives that neighbour's similarity as a vote; votes are summed per item IN NEIGHBOUR ORDER (the
// reference's defaultdict loop) and the K heaviest items are returned, (weight desc, item id asc).
#include "../../include/sss_b200.h"
#include "common.cuh"

namespace sss {

constexpr int IV_CAP = 8192;  // (item, weight) pairs per query

__device__ __forceinline__ void iv_bitonic(uint64_t* e, int P, bool descending) {
  for (int k2 = 2; k2 <= P; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t a = e[i], b = e[ixj];
          bool first_dir = (i & k2) == 0;
          bool want_desc = first_dir == descending;
          if (want_desc ? (a < b) : (a > b)) {
            e[i] = b;
            e[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
}

// one block (512 threads) per query
__global__ void __launch_bounds__(512) item_vote_kernel(const float* __restrict__ D, const int64_t* __restrict__ I, int s,
                                                        const int64_t* __restrict__ item_off,
                                                        const int64_t* __restrict__ items, int64_t n_sessions, int K,
                                                        int64_t* __restrict__ out_items, float* __restrict__ out_w,
                                                        int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char iv_smem[];
  uint64_t* key = reinterpret_cast<uint64_t*>(iv_smem);  // [IV_CAP] item << 20 | arrival, ascending
  uint64_t* agg = key + IV_CAP;                          // [IV_CAP] weight key << 32 | ~item, descending
  float* w = reinterpret_cast<float*>(agg + IV_CAP);     // [IV_CAP] weight by arrival
  int* start = reinterpret_cast<int*>(w + IV_CAP);       // [s + 1] exclusive prefix of neighbour list lengths
  __shared__ int s_total;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {  // s <= a few hundred: a serial prefix is fine
    int tot = 0;
    for (int j = 0; j < s; ++j) {
      start[j] = tot;
      const int64_t sess = I[(size_t)q * s + j];
      if (sess >= 0 && sess < n_sessions) tot += (int)(item_off[sess + 1] - item_off[sess]);
    }
    start[s] = tot;
    s_total = tot;
  }
  __syncthreads();
  int total = s_total;
  if (total > IV_CAP) {
    if (tid == 0) *overflow = 1;
    total = IV_CAP;
  }
  int P = 2;
  while (P < total) P <<= 1;
  for (int j = tid; j < s; j += blockDim.x) {
    const int64_t sess = I[(size_t)q * s + j];
    if (sess < 0 || sess >= n_sessions) continue;
    const float wj = D[(size_t)q * s + j];
    int pos = start[j];
    for (int64_t t = item_off[sess]; t < item_off[sess + 1] && pos < total; ++t, ++pos) {
      key[pos] = ((uint64_t)items[t] << 20) | (uint64_t)pos;
      w[pos] = wj;
    }
  }
  for (int i = total + tid; i < P; i += blockDim.x) key[i] = ~0ull;
  __syncthreads();
  iv_bitonic(key, P, false);
  for (int i = tid; i < P; i += blockDim.x) {
    uint64_t out = 0ull;
    if (i < total) {
      const uint64_t it = key[i] >> 20;
      if (i == 0 || (key[i - 1] >> 20) != it) {  // head of the item's run: add its votes in arrival order
        float acc = 0.0f;
        for (int r = i; r < total && (key[r] >> 20) == it; ++r) acc += w[key[r] & 0xFFFFFu];
        out = ((uint64_t)score_key(acc) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)it);
      }
    }
    agg[i] = out;
  }
  __syncthreads();
  iv_bitonic(agg, P, true);
  for (int j = tid; j < K; j += blockDim.x) {
    const uint64_t a = j < P ? agg[j] : 0ull;
    out_items[(size_t)q * K + j] = a ? (int64_t)(0xFFFFFFFFu - (uint32_t)a) : -1;
    out_w[(size_t)q * K + j] = a ? key_score((uint32_t)(a >> 32)) : 0.0f;
  }
}

}  // namespace sss

using namespace sss;

extern "C" int sss_item_vote(const float* D, const int64_t* I, int64_t nq, int s, const int64_t* item_off,
                             const int64_t* items, int64_t n_sessions, int K, int64_t* out_items, float* out_w, int device,
                             void* stream) {
  SSS_REQUIRE(D && I && item_off && items && out_items && out_w, "sss_item_vote: NULL buffer");
  SSS_REQUIRE(nq >= 0 && s >= 1 && s <= 4096 && K >= 1 && K <= IV_CAP, "sss_item_vote: bad shape");
  if (nq == 0) return 0;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaStream_t st = (cudaStream_t)stream;
  int* flag = nullptr;
  int host_flag = 0;
  cudaError_t err = cudaMalloc((void**)&flag, sizeof(int));
  if (err == cudaSuccess) err = cudaMemsetAsync(flag, 0, sizeof(int), st);
  const size_t smem = (size_t)IV_CAP * (8 + 8 + 4) + (size_t)(s + 1) * 4;
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(item_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err == cudaSuccess) {
    item_vote_kernel<<<(unsigned)nq, 512, smem, st>>>(D, I, s, item_off, items, n_sessions, K, out_items, out_w, flag);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (flag) cudaFree(flag);
  cudaSetDevice(prev);
  if (err != cudaSuccess) {
    set_error(std::string("sss_item_vote: ") + cudaGetErrorString(err));
    return 1;
  }
  SSS_REQUIRE(host_flag == 0, "sss_item_vote: more than 8192 (item, vote) pairs for one query");
  return 0;
}
