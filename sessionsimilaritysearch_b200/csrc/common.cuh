// common.cuh — shared device/host helpers for libsss_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace sss {

// ---- error plumbing (C ABI: int status + thread-local message) --------------------------------------
void set_error(const std::string& msg);
#define SSS_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::sss::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                       ":" + std::to_string(__LINE__) + ")");                                     \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)
#define SSS_REQUIRE(cond, msg)     \
  do {                             \
    if (!(cond)) {                 \
      ::sss::set_error(msg);       \
      return 1;                    \
    }                              \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: every launcher remembers what it has set
// per device (a process may hold handles on several GPUs), under a lock (handles on different threads).
struct SmemAttr {
  int set[64] = {0};
  template <typename F>
  int ensure(F* fn, int bytes) {
    int dev = 0;
    SSS_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 63;
    int cur = __atomic_load_n(&set[dev], __ATOMIC_ACQUIRE);
    if (cur >= bytes) return 0;
    SSS_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    // (two racing threads both set an attribute that only ever grows; the larger value wins below)
    while (cur < bytes && !__atomic_compare_exchange_n(&set[dev], &cur, bytes, false, __ATOMIC_RELEASE, __ATOMIC_ACQUIRE)) {
    }
    return 0;
  }
};

// SM count of the current device (cached per device)
inline int current_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = __atomic_load_n(&cache[dev], __ATOMIC_ACQUIRE);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    __atomic_store_n(&cache[dev], n, __ATOMIC_RELEASE);
  }
  return n;
}

// ---- candidate encoding -----------------------------------------------------------------------------
// A candidate is one uint64: (order-preserving key of the fp32 score) << 32 | (0xFFFFFFFF - id).
// Sorting these descending gives (score desc, id asc) — the order rule of the whole library.
// 0 is never a valid candidate (key(-inf) = 0x007FFFFF > 0), so 0 marks an empty slot.
__host__ __device__ __forceinline__ uint32_t score_key(float f) {
  f = f + 0.0f;  // -0.0 -> +0.0 so that equal floats have equal keys
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_score(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t pack_cand(uint32_t key, uint32_t id) {
  return ((uint64_t)key << 32) | (uint64_t)(0xFFFFFFFFu - id);
}
__host__ __device__ __forceinline__ uint32_t cand_key(uint64_t c) { return (uint32_t)(c >> 32); }
__host__ __device__ __forceinline__ uint32_t cand_id(uint64_t c) { return 0xFFFFFFFFu - (uint32_t)c; }

// ---- scan <-> select shared state ------------------------------------------------------------------
// One of these per search call, all device pointers.
struct SelectState {
  float* thr;        // [nq_pad] a row can only matter if score > thr (strict); +inf for padding queries
  uint32_t* cnt;     // [nq_pad] candidates appended so far (may exceed cap -> overflow)
  uint32_t* nret;    // [nq_pad] entries [0, nret) are retained (already reduced) from earlier waves
  uint64_t* cand;    // [nq_pad, cap]
  float* margin;     // [nq_pad] filter slack of the bf16 scan in EXACT mode, else 0
  float* qn2;        // [nq_pad] ||q||^2 (L2 metric on the tensor path: -dist = 2 * tensor score - ||q||^2)
  uint32_t* skip_list;  // [nq_pad] queries the small refine left to the large one (this wave)
  uint32_t* skip_cnt;   // [2] length of skip_list, indexed by wave parity
  int* overflow;     // [1] set when any list / record region overflowed
  int cap;
};

// Tensor-core scan hit record: one epilogue lane saw max(32 consecutive scores) > thr and dumped them.
// Records of query q written by scan CTA x (blockIdx.x) and epilogue warpgroup w live in the private
// sub-region ((q * grid_x + x) * 2 + w) of kRecSubCap records: no atomics on the producer side, and refine
// finds a query's records without a scatter pass.
struct __align__(16) HitRecord {
  uint32_t q;         // query index (padded space)
  uint32_t row_base;  // DB row of v[0]
  uint32_t pad0, pad1;
  float v[32];
};
static_assert(sizeof(HitRecord) == 144, "HitRecord must be 9 x 16 bytes");

constexpr int kRecSubCap = 16;   // records per (query, scan CTA, epilogue warpgroup) and wave
constexpr int kTileRows = 128;   // DB rows per tensor-core tile (UMMA N)
constexpr int kTileQ = 128;      // queries per m-tile (UMMA M)

}  // namespace sss
