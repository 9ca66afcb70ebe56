// encoder.cu — the session encoder forward after the text embedder:
//   UnifyPoolingGraphLevelEncoder.forward   model/model.py:279-351 (use_id_embedding=False, eval)
//   HeteroGGNN.forward                      model/gnn.py:64-81
//   GATConv / GatedGraphConv / HeteroConv   torch_geometric 2.0.4 [recalled], SURVEY.md Appendix A
//   PositionalAttentionPooling.forward      model/gnn.py:193-217
//   BinarizeHead.forward (eval, mlp=None)   model/model.py:117-138
//
// The dense linears are plain fp32 GEMMs and go to cuBLAS (pedantic fp32, no TF32); everything that is not a
// plain GEMM is hand written: destination-sorted CSR build (with GATConv's bipartite self-loop quirk), fused
// edge-softmax + weighted segment sum (GAT), segment sum (GatedGraphConv), GRU gates + residual relu, the
// pooling's expand / positional concat / tanh, per-graph means and the gated attention reduction.  All
// segment reductions walk their rows in a fixed order, so a forward is deterministic run to run.
#include <dlfcn.h>
#include <math.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/sss_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace sss {

// ---- cuBLAS through dlopen: the search path of the library must not depend on it ---------------------
typedef void* cublasHandle_t;
typedef int (*cublasCreate_t)(cublasHandle_t*);
typedef int (*cublasDestroy_t)(cublasHandle_t);
typedef int (*cublasSetStream_t)(cublasHandle_t, cudaStream_t);
typedef int (*cublasSetMathMode_t)(cublasHandle_t, int);
typedef int (*cublasSetEmulationStrategy_t)(cublasHandle_t, int);
typedef int (*cublasSgemm_t)(cublasHandle_t, int, int, int, int, int, const float*, const float*, int, const float*,
                             int, const float*, float*, int);
struct Cublas {
  void* lib = nullptr;
  cublasCreate_t create = nullptr;
  cublasDestroy_t destroy = nullptr;
  cublasSetStream_t set_stream = nullptr;
  cublasSetMathMode_t set_math = nullptr;
  cublasSgemm_t sgemm = nullptr;
  cublasSetEmulationStrategy_t set_emulation = nullptr;  // cuBLAS >= 12.9 only
  bool load() {
    if (lib) return true;
    const char* names[] = {"libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12", "libcublas.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    create = (cublasCreate_t)dlsym(lib, "cublasCreate_v2");
    destroy = (cublasDestroy_t)dlsym(lib, "cublasDestroy_v2");
    set_stream = (cublasSetStream_t)dlsym(lib, "cublasSetStream_v2");
    set_math = (cublasSetMathMode_t)dlsym(lib, "cublasSetMathMode");
    sgemm = (cublasSgemm_t)dlsym(lib, "cublasSgemm_v2");
    set_emulation = (cublasSetEmulationStrategy_t)dlsym(lib, "cublasSetEmulationStrategy");
    return create && destroy && set_stream && set_math && sgemm;
  }
};
static Cublas g_cublas;
constexpr int kOpN = 0, kOpT = 1, kPedanticMath = 2, kBf16x9Math = 4, kEmulationEager = 2;

// row-major C[M,N] (ldc) = A[M,K] (lda) * B^T, B row-major [N,K] (ldb)      (x @ W.T, torch nn.Linear)
static int gemm_nt(cublasHandle_t h, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                   int ldc) {
  if (M == 0 || N == 0) return 0;
  const float one = 1.0f, zero = 0.0f;
  int rc = g_cublas.sgemm(h, kOpT, kOpN, N, M, K, &one, B, ldb, A, lda, &zero, C, ldc);
  SSS_REQUIRE(rc == 0, "cublasSgemm failed with status " + std::to_string(rc));
  return 0;
}
// row-major C[M,N] = A[M,K] * B, B row-major [K,N] (ldb)                      (x @ W, GatedGraphConv)
static int gemm_nn(cublasHandle_t h, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                   int ldc) {
  if (M == 0 || N == 0) return 0;
  const float one = 1.0f, zero = 0.0f;
  int rc = g_cublas.sgemm(h, kOpN, kOpN, N, M, K, &one, B, ldb, A, lda, &zero, C, ldc);
  SSS_REQUIRE(rc == 0, "cublasSgemm failed with status " + std::to_string(rc));
  return 0;
}

// ---- CSR by destination ----------------------------------------------------------------------------------
// GATConv(add_self_loops=True) on a bipartite edge set: drop edges with src == dst (batch-global indices), then
// append (i, i) for i < n_loop = min(N_src, N_dst).  GatedGraphConv: n_loop = 0, nothing dropped.
__global__ void csr_count_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                 int drop_self, int n_loop, int* __restrict__ deg) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < E) {
    if (!(drop_self && src[t] == dst[t])) atomicAdd(&deg[dst[t]], 1);
  } else if (t < E + n_loop) {
    atomicAdd(&deg[t - E], 1);
  }
}
// single block exclusive scan (n is a few thousand)
__global__ void csr_scan_kernel(const int* __restrict__ deg, int n, int* __restrict__ rowptr, int* __restrict__ cursor) {
  __shared__ int carry;
  __shared__ int buf[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < n ? deg[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    int excl = carry + buf[threadIdx.x] - v;
    if (i < n) {
      rowptr[i] = excl;
      cursor[i] = excl;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowptr[n] = carry;
}
__global__ void csr_fill_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                int drop_self, int n_loop, int* __restrict__ cursor, int* __restrict__ col) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < E) {
    if (!(drop_self && src[t] == dst[t])) col[atomicAdd(&cursor[dst[t]], 1)] = (int)src[t];
  } else if (t < E + n_loop) {
    int i = (int)(t - E);
    col[atomicAdd(&cursor[i], 1)] = i;
  }
}
// lists are tiny: insertion sort by source makes the summation order deterministic
__global__ void csr_sort_kernel(const int* __restrict__ rowptr, int n, int* __restrict__ col) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int b = rowptr[i], e = rowptr[i + 1];
  for (int a = b + 1; a < e; ++a) {
    int v = col[a], j = a - 1;
    while (j >= b && col[j] > v) {
      col[j + 1] = col[j];
      --j;
    }
    col[j + 1] = v;
  }
}

// ---- per-node attention scalars: a[i] = <X[i, :], att> ---------------------------------------------------
__global__ void rowdot_kernel(const float* __restrict__ X, int ld, int n, int H, const float* __restrict__ att,
                              float* __restrict__ out) {
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x & 31;
  if (row >= n) return;
  float acc = 0.0f;
  for (int c = lane; c < H; c += 32) acc = fmaf(X[(size_t)row * ld + c], att[c], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc;
}

// ---- GAT: edge softmax + weighted segment sum, one block per destination ---------------------------------
// out[i, :] = sum_k alpha_k * S[j_k, :] + bias, alpha = softmax_k(leaky_relu(a_s[j_k] + a_d[i], 0.2)) with the
// PyG denominator (sum + 1e-16); a destination without incoming edges gets the bias.  relu_out: write relu(out).
__global__ void gat_aggregate_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                     const float* __restrict__ S, int lds, const float* __restrict__ a_s,
                                     const float* __restrict__ a_d, const float* __restrict__ bias, int H,
                                     float* __restrict__ out, int ldo, int relu_out) {
  const int i = blockIdx.x;
  const int b = rowptr[i], e = rowptr[i + 1];
  const float ad = a_d[i];
  float mx = -INFINITY;
  for (int k = b; k < e; ++k) {
    float v = a_s[col[k]] + ad;
    v = v > 0.0f ? v : 0.2f * v;
    mx = fmaxf(mx, v);
  }
  float den = 0.0f;
  for (int k = b; k < e; ++k) {
    float v = a_s[col[k]] + ad;
    v = v > 0.0f ? v : 0.2f * v;
    den += expf(v - mx);
  }
  den += 1e-16f;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.0f;
    for (int k = b; k < e; ++k) {
      const int j = col[k];
      float v = a_s[j] + ad;
      v = v > 0.0f ? v : 0.2f * v;
      const float alpha = expf(v - mx) / den;
      acc += S[(size_t)j * lds + c] * alpha;
    }
    acc += bias[c];
    out[(size_t)i * ldo + c] = relu_out ? fmaxf(acc, 0.0f) : acc;
  }
}

// ---- GatedGraphConv: A[i, :] = sum_k M[j_k, :] ---------------------------------------------------------
__global__ void segsum_rows_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                   const float* __restrict__ M, int ldm, int H, float* __restrict__ out, int ldo) {
  const int i = blockIdx.x;
  const int b = rowptr[i], e = rowptr[i + 1];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.0f;
    for (int k = b; k < e; ++k) acc += M[(size_t)col[k] * ldm + c];
    out[(size_t)i * ldo + c] = acc;
  }
}

// ---- GRUCell gates + HeteroConv sum + relu ----------------------------------------------------------------
// r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h = (1 - z) * n + z * xpad
// (xpad = x zero padded to H); X_next = relu(Gp + h).  gi / gh arrive without their biases.
__global__ void gru_relu_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                const float* __restrict__ X, int ldx, int in_w, const float* __restrict__ Gp, int n,
                                int H, float* __restrict__ out, int ldo) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * H) return;
  int row = (int)(t / H), c = (int)(t % H);
  const float* gir = gi + (size_t)row * 3 * H;
  const float* ghr = gh + (size_t)row * 3 * H;
  float ir = gir[c] + b_ih[c], iz = gir[H + c] + b_ih[H + c], in_ = gir[2 * H + c] + b_ih[2 * H + c];
  float hr = ghr[c] + b_hh[c], hz = ghr[H + c] + b_hh[H + c], hn = ghr[2 * H + c] + b_hh[2 * H + c];
  float r = 1.0f / (1.0f + expf(-(ir + hr)));
  float z = 1.0f / (1.0f + expf(-(iz + hz)));
  float nn = tanhf(in_ + r * hn);
  float x = c < in_w ? X[(size_t)row * ldx + c] : 0.0f;
  float h = (1.0f - z) * nn + z * x;
  out[(size_t)row * ldo + c] = fmaxf(Gp[(size_t)row * H + c] + h, 0.0f);
}

// ---- pooling ---------------------------------------------------------------------------------------------
// occurrence -> product map from cnt (exclusive scan done with the CSR scan kernel): occ_of[prefix[p] + t] = p
__global__ void expand_map_kernel(const int* __restrict__ prefix, int n_p, int* __restrict__ occ_prod) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_p) return;
  for (int t = prefix[p]; t < prefix[p + 1]; ++t) occ_prod[t] = p;
}
__global__ void cnt_to_int_kernel(const int64_t* __restrict__ cnt, int n, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)cnt[i];
}
// U[r, :] = tanh([lin[node(r), :] + b | PE[pos(r), :]]); rows [0, n_e) are product occurrences, then queries
__global__ void pool_nodes_kernel(const float* __restrict__ up_lin, const float* __restrict__ uq_lin,
                                  const float* __restrict__ bp, const float* __restrict__ bq,
                                  const float* __restrict__ pe, const int* __restrict__ occ_prod,
                                  const int64_t* __restrict__ product_pos, const int64_t* __restrict__ query_pos,
                                  const int64_t* __restrict__ product_batch, const int64_t* __restrict__ query_batch,
                                  int n_e, int n_q, int lin_w, int msl, float* __restrict__ U,
                                  int* __restrict__ node_graph) {
  const int r = blockIdx.x;
  const int W = lin_w + msl;
  const bool is_prod = r < n_e;
  const int node = is_prod ? occ_prod[r] : r - n_e;
  const float* lin = (is_prod ? up_lin : uq_lin) + (size_t)node * lin_w;
  const float* b = is_prod ? bp : bq;
  const int64_t pos = is_prod ? product_pos[r] : query_pos[node];
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float v = c < lin_w ? lin[c] + b[c] : pe[pos * msl + (c - lin_w)];
    U[(size_t)r * W + c] = tanhf(v);
  }
  if (threadIdx.x == 0) node_graph[r] = (int)(is_prod ? product_batch[node] : query_batch[node]);
}
// graph -> [first, last) rows among the product occurrences and among the queries (both batch vectors are sorted)
__global__ void graph_ranges_kernel(const int* __restrict__ node_graph, int n_e, int n_tot, int* __restrict__ ranges) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_tot) return;
  const int g = node_graph[r];
  const int part = r < n_e ? 0 : 1;
  const bool first = (r == 0) || (r == n_e) || node_graph[r - 1] != g;
  const bool last = (r == n_e - 1) || (r == n_tot - 1) || node_graph[r + 1] != g;
  if (first) ranges[g * 4 + part * 2] = r;
  if (last) ranges[g * 4 + part * 2 + 1] = r + 1;
}
// mean over the rows of a graph, optionally weighted per row: out[g, :] = sum_r w[r] * U[r, :] / count
__global__ void graph_mean_kernel(const float* __restrict__ U, int W, const int* __restrict__ ranges,
                                  const float* __restrict__ w, float* __restrict__ out) {
  const int g = blockIdx.x;
  const int p0 = ranges[g * 4], p1 = ranges[g * 4 + 1], q0 = ranges[g * 4 + 2], q1 = ranges[g * 4 + 3];
  const float cnt = fmaxf((float)((p1 - p0) + (q1 - q0)), 1.0f);
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float acc = 0.0f;
    for (int r = p0; r < p1; ++r) acc += w ? U[(size_t)r * W + c] * w[r] : U[(size_t)r * W + c];
    for (int r = q0; r < q1; ++r) acc += w ? U[(size_t)r * W + c] * w[r] : U[(size_t)r * W + c];
    out[(size_t)g * W + c] = acc / cnt;
  }
}
// att[r] = sum_c w_att[c] * sigmoid(A[r, c] + b_n[c] + Bc[graph(r), c])
__global__ void pool_att_kernel(const float* __restrict__ A, const float* __restrict__ bn, const float* __restrict__ Bc,
                                const int* __restrict__ node_graph, const float* __restrict__ w_att, int W,
                                float* __restrict__ att) {
  const int r = blockIdx.x;
  const float* a = A + (size_t)r * W;
  const float* bc = Bc + (size_t)node_graph[r] * W;
  float acc = 0.0f;
  for (int c = threadIdx.x; c < W; c += blockDim.x)
    acc += w_att[c] * (1.0f / (1.0f + expf(-(a[c] + bn[c] + bc[c]))));
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) att[r] = v;
  }
}

__global__ void nan_flag_kernel(const float* __restrict__ x, int64_t n, int32_t* __restrict__ flag) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && isnan(x[t])) *flag = 1;
}
__global__ void copy_cols_kernel(const float* __restrict__ src, int n, int w, float* __restrict__ dst, int ldd) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * w) return;
  dst[(size_t)(t / w) * ldd + (t % w)] = src[t];
}
// BinarizeHead eval: sign(v + b) numerically ((sign - tanh).detach() + tanh), model/model.py:137
__global__ void sign_bias_kernel(float* __restrict__ x, const float* __restrict__ b, int64_t n, int w) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * w) return;
  float v = x[t] + b[t % w];
  float s = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);
  float th = tanhf(v);
  x[t] = (s - th) + th;
}

}  // namespace sss

using namespace sss;

struct sss_encoder {
  int device = 0;
  sss_encoder_shape_t sh{};
  std::map<std::string, float*> params;
  std::map<std::string, int64_t> numel;
  cublasHandle_t blas = nullptr;
  int math = SSS_ENCODER_MATH_FP32;
  // split-bf16 copies for the tensor-core GEMM (SSS_ENCODER_MATH_BF16X3): weights once per (pointer, shape),
  // activations once per forward (several linears read the same node features)
  struct Split {
    const float* src; int rows, cols; int64_t ld; int transposed;
    void* hi; void* lo; int rows_pad, cols_pad;
  };
  std::vector<Split> w_split;   // persistent (cudaMalloc), dropped when a parameter is replaced
  std::vector<Split> a_split;   // arena, per forward
  int* gemm_flag = nullptr;     // watchdog code of the tensor-core GEMM
  // Workspace arena: slabs are bump-allocated per forward call and kept across calls (cudaMalloc/cudaFree per
  // buffer cost more than the whole forward).  A call that needed more than one slab is followed by one
  // consolidation at the start of the next call; `done` orders reuse across streams.
  std::vector<std::pair<char*, size_t>> slabs;
  size_t used = 0;        // bytes taken from the last slab
  size_t requested = 0;   // bytes requested by the current / last call
  cudaEvent_t done = nullptr;
  bool done_recorded = false;
  ~sss_encoder() {}
};

namespace {
constexpr size_t kSlabAlign = 256, kMinSlab = (size_t)32 << 20;
int ws_begin(sss_encoder* e, cudaStream_t st) {
  if (!e->done) SSS_CUDA_OK(cudaEventCreateWithFlags(&e->done, cudaEventDisableTiming));
  if (e->slabs.size() > 1) {  // the last call outgrew its slab: one slab of the full size from now on
    if (e->done_recorded) SSS_CUDA_OK(cudaEventSynchronize(e->done));
    for (auto& s : e->slabs) cudaFree(s.first);
    e->slabs.clear();
    const size_t want = e->requested + e->requested / 4;
    char* v = nullptr;
    SSS_CUDA_OK(cudaMalloc((void**)&v, want));
    e->slabs.emplace_back(v, want);
  }
  if (e->done_recorded) SSS_CUDA_OK(cudaStreamWaitEvent(st, e->done, 0));
  e->used = 0;
  e->requested = 0;
  e->a_split.clear();
  return 0;
}
int ws_end(sss_encoder* e, cudaStream_t st) {
  SSS_CUDA_OK(cudaEventRecord(e->done, st));
  e->done_recorded = true;
  return 0;
}
template <typename T>
int ws_alloc(sss_encoder* e, T** p, size_t count) {
  size_t bytes = ((count > 0 ? count : 1) * sizeof(T) + kSlabAlign - 1) / kSlabAlign * kSlabAlign;
  e->requested += bytes;
  if (e->slabs.empty() || e->used + bytes > e->slabs.back().second) {
    const size_t want = bytes > kMinSlab ? bytes : kMinSlab;
    char* v = nullptr;
    SSS_CUDA_OK(cudaMalloc((void**)&v, want));
    e->slabs.emplace_back(v, want);
    e->used = 0;
  }
  *p = (T*)(e->slabs.back().first + e->used);
  e->used += bytes;
  return 0;
}
void ws_release(sss_encoder* e) {
  if (e->done_recorded) cudaEventSynchronize(e->done);
  for (auto& s : e->slabs) cudaFree(s.first);
  e->slabs.clear();
  if (e->done) cudaEventDestroy(e->done);
  e->done = nullptr;
  e->done_recorded = false;
}
void drop_weight_splits(sss_encoder* e) {
  for (auto& s : e->w_split) {
    cudaFree(s.hi);
    cudaFree(s.lo);
  }
  e->w_split.clear();
}

// C[M,N] (ldc) = A[M,K] (lda) * op(B): b_transposed == 0: B is [N,K] row-major (ldb) -> x @ W.T;
//                                      b_transposed == 1: B is [K,N] row-major (ldb) -> x @ W
int enc_gemm(sss_encoder* e, cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
             int b_transposed, float* C, int ldc) {
  if (M == 0 || N == 0) return 0;
  if (e->math != SSS_ENCODER_MATH_BF16X3)
    return b_transposed ? gemm_nn(e->blas, M, N, K, A, lda, B, ldb, C, ldc) : gemm_nt(e->blas, M, N, K, A, lda, B, ldb, C, ldc);
  const int k_pad = (K + 63) / 64 * 64;
  auto find = [](std::vector<sss_encoder::Split>& v, const float* src, int rows, int cols, int64_t ld, int tr) {
    for (auto& s : v)
      if (s.src == src && s.rows == rows && s.cols == cols && s.ld == ld && s.transposed == tr) return &s;
    return (sss_encoder::Split*)nullptr;
  };
  sss_encoder::Split* ws = find(e->w_split, B, N, K, ldb, b_transposed);
  if (!ws) {
    sss_encoder::Split s{B, N, K, (int64_t)ldb, b_transposed, nullptr, nullptr, (N + 127) / 128 * 128, k_pad};
    const size_t bytes = (size_t)s.rows_pad * s.cols_pad * 2;
    SSS_CUDA_OK(cudaMalloc(&s.hi, bytes));
    SSS_CUDA_OK(cudaMalloc(&s.lo, bytes));
    if (launch_split_bf16(B, N, K, ldb, b_transposed, s.hi, s.lo, s.rows_pad, s.cols_pad, st)) return 1;
    e->w_split.push_back(s);
    ws = &e->w_split.back();
  }
  sss_encoder::Split* as = find(e->a_split, A, M, K, lda, 0);
  if (!as) {
    sss_encoder::Split s{A, M, K, (int64_t)lda, 0, nullptr, nullptr, (M + 127) / 128 * 128, k_pad};
    uint16_t *hi, *lo;
    if (ws_alloc(e, &hi, (size_t)s.rows_pad * s.cols_pad) || ws_alloc(e, &lo, (size_t)s.rows_pad * s.cols_pad)) return 1;
    s.hi = hi;
    s.lo = lo;
    if (launch_split_bf16(A, M, K, lda, 0, s.hi, s.lo, s.rows_pad, s.cols_pad, st)) return 1;
    e->a_split.push_back(s);
    as = &e->a_split.back();
  }
  return launch_gemm_bf16x3(as->hi, as->lo, as->rows_pad, ws->hi, ws->lo, ws->rows_pad, k_pad, C, M, N, ldc,
                            e->gemm_flag, st);
}

// Several linears that read the same input, written side by side into C (column blocks of widths n[i]):
//   C[:, off_i : off_i + n_i] = A * op(B_i).  On the tensor-core path they run as ONE GEMM against the row-wise
// concatenation of the (split) weights; otherwise one library GEMM each.
struct LinPart {
  const float* B;
  int ldb, transposed, n;
};
int enc_gemm_fused(sss_encoder* e, cudaStream_t st, int M, int K, const float* A, int lda, const LinPart* parts,
                   int n_parts, float* C, int ldc) {
  int n_total = 0;
  for (int i = 0; i < n_parts; ++i) n_total += parts[i].n;
  if (e->math != SSS_ENCODER_MATH_BF16X3 || n_parts == 1) {
    int off = 0;
    for (int i = 0; i < n_parts; ++i) {
      if (enc_gemm(e, st, M, parts[i].n, K, A, lda, parts[i].B, parts[i].ldb, parts[i].transposed, C + off, ldc)) return 1;
      off += parts[i].n;
    }
    return 0;
  }
  if (M == 0 || n_total == 0) return 0;
  const int k_pad = (K + 63) / 64 * 64;
  // the fused weight is cached under the first part's pointer with the total row count (rows = n_total marks it)
  sss_encoder::Split* ws = nullptr;
  for (auto& s : e->w_split)
    if (s.src == parts[0].B && s.rows == n_total && s.cols == K && s.ld == -(int64_t)n_parts) ws = &s;
  if (!ws) {
    sss_encoder::Split s{parts[0].B, n_total, K, -(int64_t)n_parts, 0, nullptr, nullptr, (n_total + 127) / 128 * 128, k_pad};
    const size_t bytes = (size_t)s.rows_pad * s.cols_pad * 2;
    SSS_CUDA_OK(cudaMalloc(&s.hi, bytes));
    SSS_CUDA_OK(cudaMalloc(&s.lo, bytes));
    int off = 0;
    for (int i = 0; i < n_parts; ++i) {
      const int rows_out = i + 1 == n_parts ? s.rows_pad - off : parts[i].n;  // the last part also zeroes the padding
      if (launch_split_bf16(parts[i].B, parts[i].n, K, parts[i].ldb, parts[i].transposed,
                            (uint16_t*)s.hi + (size_t)off * k_pad, (uint16_t*)s.lo + (size_t)off * k_pad, rows_out, k_pad, st))
        return 1;
      off += parts[i].n;
    }
    e->w_split.push_back(s);
    ws = &e->w_split.back();
  }
  sss_encoder::Split* as = nullptr;
  for (auto& s : e->a_split)
    if (s.src == A && s.rows == M && s.cols == K && s.ld == lda) as = &s;
  if (!as) {
    sss_encoder::Split s{A, M, K, (int64_t)lda, 0, nullptr, nullptr, (M + 127) / 128 * 128, k_pad};
    uint16_t *hi, *lo;
    if (ws_alloc(e, &hi, (size_t)s.rows_pad * s.cols_pad) || ws_alloc(e, &lo, (size_t)s.rows_pad * s.cols_pad)) return 1;
    s.hi = hi;
    s.lo = lo;
    if (launch_split_bf16(A, M, K, lda, 0, s.hi, s.lo, s.rows_pad, s.cols_pad, st)) return 1;
    e->a_split.push_back(s);
    as = &e->a_split.back();
  }
  return launch_gemm_bf16x3(as->hi, as->lo, as->rows_pad, ws->hi, ws->lo, ws->rows_pad, k_pad, C, M, n_total, ldc,
                            e->gemm_flag, st);
}

struct Csr {
  int* rowptr = nullptr;
  int* col = nullptr;
};
int build_csr(sss_encoder* e, const int64_t* src, const int64_t* dst, int64_t E, int n_dst, int drop_self, int n_loop,
              Csr* out, cudaStream_t st) {
  int *deg, *cursor;
  if (ws_alloc(e, &deg, n_dst + 1) || ws_alloc(e, &cursor, n_dst + 1) || ws_alloc(e, &out->rowptr, n_dst + 1) ||
      ws_alloc(e, &out->col, E + n_loop + 1))
    return 1;
  SSS_CUDA_OK(cudaMemsetAsync(deg, 0, sizeof(int) * (n_dst + 1), st));
  int64_t tot = E + n_loop;
  if (tot > 0) csr_count_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, dst, E, drop_self, n_loop, deg);
  csr_scan_kernel<<<1, 1024, 0, st>>>(deg, n_dst, out->rowptr, cursor);
  if (tot > 0) csr_fill_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, dst, E, drop_self, n_loop, cursor, out->col);
  if (n_dst > 0) csr_sort_kernel<<<(n_dst + 127) / 128, 128, 0, st>>>(out->rowptr, n_dst, out->col);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}
const float* P(sss_encoder* e, const std::string& k, int64_t expect) {
  auto it = e->params.find(k);
  if (it == e->params.end()) {
    set_error("encoder parameter missing: " + k);
    return nullptr;
  }
  if (e->numel[k] != expect) {
    set_error("encoder parameter " + k + " has " + std::to_string(e->numel[k]) + " elements, expected " +
              std::to_string(expect));
    return nullptr;
  }
  return it->second;
}
}  // namespace

extern "C" int sss_encoder_create(sss_encoder_t** out, int device, const sss_encoder_shape_t* shape) {
  SSS_REQUIRE(out && shape, "sss_encoder_create: NULL argument");
  SSS_REQUIRE(shape->in_dim >= 1 && shape->n_layers >= 1 &&
                  shape->out_dim > shape->max_seq_len && shape->max_seq_len >= 1,
              "sss_encoder_create: bad shape");
  SSS_REQUIRE(shape->in_dim <= shape->hidden,
              "The number of input channels is not allowed to be larger than the number of output channels");
  int ndev = 0;
  SSS_CUDA_OK(cudaGetDeviceCount(&ndev));
  SSS_REQUIRE(device >= 0 && device < ndev, "sss_encoder_create: no such CUDA device");
  SSS_REQUIRE(g_cublas.load(), "cuBLAS (libcublas.so.12) could not be loaded for the encoder's dense linears");
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  sss_encoder* e = new sss_encoder();
  e->device = device;
  e->sh = *shape;
  int rc = g_cublas.create(&e->blas);
  if (rc == 0) rc = g_cublas.set_math(e->blas, kPedanticMath);
  cudaSetDevice(prev);
  if (rc != 0) {
    delete e;
    set_error("cublasCreate failed with status " + std::to_string(rc));
    return 1;
  }
  *out = e;
  return 0;
}

extern "C" int sss_encoder_set_math(sss_encoder_t* e, int math) {
  SSS_REQUIRE(e != nullptr, "sss_encoder_set_math: NULL encoder");
  SSS_REQUIRE(math == SSS_ENCODER_MATH_FP32 || math == SSS_ENCODER_MATH_BF16X9 || math == SSS_ENCODER_MATH_BF16X3,
              "sss_encoder_set_math: unknown mode");
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(e->device);
  int rc;
  if (math == SSS_ENCODER_MATH_BF16X9) {
    rc = g_cublas.set_emulation ? g_cublas.set_math(e->blas, kBf16x9Math) : 1;
    if (rc == 0) rc = g_cublas.set_emulation(e->blas, kEmulationEager);
    if (rc != 0) g_cublas.set_math(e->blas, kPedanticMath);
  } else {
    rc = g_cublas.set_math(e->blas, kPedanticMath);
    if (rc == 0 && math == SSS_ENCODER_MATH_BF16X3 && !e->gemm_flag) {
      rc = cudaMalloc((void**)&e->gemm_flag, sizeof(int)) == cudaSuccess ? 0 : 1;
      if (rc == 0) rc = cudaMemset(e->gemm_flag, 0, sizeof(int)) == cudaSuccess ? 0 : 1;
    }
  }
  cudaSetDevice(prev);
  SSS_REQUIRE(rc == 0, "the loaded cuBLAS does not offer fp32 emulation on bf16 tensor cores (BF16x9 needs cuBLAS >= 12.9)");
  e->math = math;
  return 0;
}

extern "C" int sss_encoder_get_math(const sss_encoder_t* e) { return e ? e->math : -1; }

extern "C" int sss_encoder_destroy(sss_encoder_t* e) {
  if (!e) return 0;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(e->device);
  for (auto& kv : e->params) cudaFree(kv.second);
  drop_weight_splits(e);
  if (e->gemm_flag) cudaFree(e->gemm_flag);
  ws_release(e);
  if (e->blas) g_cublas.destroy(e->blas);
  cudaSetDevice(prev);
  delete e;
  return 0;
}

extern "C" int sss_encoder_set_param(sss_encoder_t* e, const char* name, const float* data, int64_t numel, int on_device,
                                     void* stream) {
  SSS_REQUIRE(e && name && data && numel > 0, "sss_encoder_set_param: bad argument");
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(e->device);
  std::string k(name);
  drop_weight_splits(e);  // split-bf16 copies are keyed by pointer: a replaced parameter may reuse an address
  auto it = e->params.find(k);
  if (it != e->params.end()) {
    cudaFree(it->second);
    e->params.erase(it);
  }
  float* d = nullptr;
  cudaError_t err = cudaMalloc((void**)&d, sizeof(float) * numel);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(d, data, sizeof(float) * numel, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                          (cudaStream_t)stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize((cudaStream_t)stream);
  cudaSetDevice(prev);
  if (err != cudaSuccess) {
    set_error(std::string("sss_encoder_set_param: ") + cudaGetErrorString(err));
    return 1;
  }
  e->params[k] = d;
  e->numel[k] = numel;
  return 0;
}

extern "C" int sss_encoder_forward(sss_encoder_t* e, const sss_graph_batch_t* bt, float* out, int32_t* nonfinite,
                                   void* stream) {
  sss_encoder_io_t io;
  io.out = out;
  io.z_query = nullptr;
  io.z_product = nullptr;
  io.run_gnn = 1;
  io.run_pooling = 1;
  io.nonfinite = nonfinite;
  return sss_encoder_forward_ex(e, bt, &io, stream);
}

extern "C" int sss_encoder_forward_ex(sss_encoder_t* e, const sss_graph_batch_t* bt, const sss_encoder_io_t* io,
                                      void* stream) {
  SSS_REQUIRE(e && bt && io, "sss_encoder_forward: NULL argument");
  float* out = io->out;
  int32_t* nonfinite = io->nonfinite;
  const bool run_gnn = io->run_gnn != 0, run_pool = io->run_pooling != 0;
  SSS_REQUIRE(run_gnn || run_pool, "sss_encoder_forward_ex: nothing to run");
  SSS_REQUIRE(!run_pool || out != nullptr, "sss_encoder_forward_ex: the pooling stage needs `out`");
  SSS_REQUIRE(run_gnn || (io->z_query && io->z_product), "sss_encoder_forward_ex: pooling alone needs z_query / z_product");
  SSS_REQUIRE(run_pool || (io->z_query && io->z_product), "sss_encoder_forward_ex: the GNN stage alone needs z_query / z_product");
  const int IN = e->sh.in_dim, H = e->sh.hidden, L = e->sh.n_layers, OUT = e->sh.out_dim, MSL = e->sh.max_seq_len;
  const int LIN = OUT - MSL, ZD = IN + L * H;
  const int B = (int)bt->n_graphs, NQ = (int)bt->n_query, NP = (int)bt->n_product, NE = (int)bt->n_expanded;
  SSS_REQUIRE(B >= 1 && NQ >= 1 && NP >= 1 && NE >= NP, "sss_encoder_forward: empty batch");
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(e->device);
  struct Restore {
    int d;
    sss_encoder* e;
    ~Restore() {
      cudaSetDevice(d);
    }
  } restore{prev, e};
  cudaStream_t st = (cudaStream_t)stream;
  SSS_REQUIRE(g_cublas.set_stream(e->blas, st) == 0, "cublasSetStream failed");
  if (ws_begin(e, st)) return 1;

  // ---- workspace
  float *Zq = io->z_query, *Zp = io->z_product;  // node embeddings [N, in + layers * hidden]: the caller's or ours
  float *Sq, *Sp, *as_q, *ad_q, *as_p, *ad_p, *Gp, *Agg, *gi, *gh, *uq_lin, *up_lin, *U, *coarse, *Aatt, *Bc, *att;
  int *cnt_i, *cnt_pre, *cursor_tmp, *occ_prod, *node_graph, *ranges;
  const int NT = NE + NQ;
  if ((!Zq && ws_alloc(e, &Zq, (size_t)NQ * ZD)) || (!Zp && ws_alloc(e, &Zp, (size_t)NP * ZD)) || ws_alloc(e, &Sq, (size_t)NQ * 2 * H) ||
      ws_alloc(e, &Sp, (size_t)NP * 3 * H) || ws_alloc(e, &as_q, NQ) || ws_alloc(e, &ad_q, NQ) || ws_alloc(e, &as_p, NP) ||
      ws_alloc(e, &ad_p, NP) || ws_alloc(e, &Gp, (size_t)NP * H) || ws_alloc(e, &Agg, (size_t)NP * H) ||
      ws_alloc(e, &gi, (size_t)NP * 3 * H) || ws_alloc(e, &gh, (size_t)NP * 3 * H) ||
      ws_alloc(e, &uq_lin, (size_t)NQ * LIN) || ws_alloc(e, &up_lin, (size_t)NP * LIN) ||
      ws_alloc(e, &U, (size_t)NT * OUT) || ws_alloc(e, &coarse, (size_t)B * OUT) || ws_alloc(e, &Aatt, (size_t)NT * OUT) ||
      ws_alloc(e, &Bc, (size_t)B * OUT) || ws_alloc(e, &att, NT) || ws_alloc(e, &cnt_i, NP + 1) ||
      ws_alloc(e, &cnt_pre, NP + 1) || ws_alloc(e, &cursor_tmp, NP + 1) || ws_alloc(e, &occ_prod, NE + 1) ||
      ws_alloc(e, &node_graph, NT) || ws_alloc(e, &ranges, (size_t)B * 4))
    return 1;
  if (nonfinite) SSS_CUDA_OK(cudaMemsetAsync(nonfinite, 0, sizeof(int32_t), st));
  if (nonfinite && run_gnn) {
    nan_flag_kernel<<<(unsigned)(((int64_t)NQ * IN + 255) / 256), 256, 0, st>>>(bt->x_query, (int64_t)NQ * IN, nonfinite);
    nan_flag_kernel<<<(unsigned)(((int64_t)NP * IN + 255) / 256), 256, 0, st>>>(bt->x_product, (int64_t)NP * IN, nonfinite);
  }
  Csr qp, pq, pp;
  if (run_gnn) {
    SSS_REQUIRE(bt->x_query && bt->x_product, "sss_encoder_forward: the GNN stage needs x_query / x_product");
    copy_cols_kernel<<<(unsigned)(((int64_t)NQ * IN + 255) / 256), 256, 0, st>>>(bt->x_query, NQ, IN, Zq, ZD);
    copy_cols_kernel<<<(unsigned)(((int64_t)NP * IN + 255) / 256), 256, 0, st>>>(bt->x_product, NP, IN, Zp, ZD);
    // ---- graph structure, shared by the three layers
    const int n_loop = NQ < NP ? NQ : NP;
    if (build_csr(e, bt->qp_src, bt->qp_dst, bt->e_qp, NP, 1, n_loop, &qp, st)) return 1;   // dst = product
    if (build_csr(e, bt->pq_src, bt->pq_dst, bt->e_pq, NQ, 1, n_loop, &pq, st)) return 1;   // dst = query
    if (build_csr(e, bt->pp_src, bt->pp_dst, bt->e_pp, NP, 0, 0, &pp, st)) return 1;        // dst = product
  }

  // ---- HeteroGGNN layers (model/gnn.py:64-81)
  for (int l = 0; run_gnn && l < L; ++l) {
    const int cin = l == 0 ? IN : H;
    e->a_split.clear();  // activation splits are keyed by pointer and buffers such as Agg are rewritten every layer
    const int off = l == 0 ? 0 : IN + (l - 1) * H;
    const int off_next = IN + l * H;
    const std::string pre = "gnn.convs." + std::to_string(l) + ".convs.";
    const std::string eqp = pre + "query__clicks__product.", epq = pre + "product__clicked by__query.",
                      epp = pre + "product__to__product.";
    const float *w_qp_src = P(e, eqp + "lin_src.weight", (int64_t)H * cin), *w_qp_dst = P(e, eqp + "lin_dst.weight", (int64_t)H * cin),
                *a_qp_src = P(e, eqp + "att_src", H), *a_qp_dst = P(e, eqp + "att_dst", H), *b_qp = P(e, eqp + "bias", H),
                *w_pq_src = P(e, epq + "lin_src.weight", (int64_t)H * cin), *w_pq_dst = P(e, epq + "lin_dst.weight", (int64_t)H * cin),
                *a_pq_src = P(e, epq + "att_src", H), *a_pq_dst = P(e, epq + "att_dst", H), *b_pq = P(e, epq + "bias", H),
                *w_g = P(e, epp + "weight", (int64_t)H * H), *w_ih = P(e, epp + "rnn.weight_ih", (int64_t)3 * H * H),
                *w_hh = P(e, epp + "rnn.weight_hh", (int64_t)3 * H * H), *b_ih = P(e, epp + "rnn.bias_ih", 3 * H),
                *b_hh = P(e, epp + "rnn.bias_hh", 3 * H);
    if (!w_qp_src || !w_qp_dst || !a_qp_src || !a_qp_dst || !b_qp || !w_pq_src || !w_pq_dst || !a_pq_src || !a_pq_dst ||
        !b_pq || !w_g || !w_ih || !w_hh || !b_ih || !b_hh)
      return 1;
    const float* Xq = Zq + off;
    const float* Xp = Zp + off;
    // query side: S_qp (source of q->p) | T_pq (destination of p->q)
    const LinPart q_parts[2] = {{w_qp_src, cin, 0, H}, {w_pq_dst, cin, 0, H}};
    if (enc_gemm_fused(e, st, NQ, cin, Xq, ZD, q_parts, 2, Sq, 2 * H)) return 1;
    // product side: T_qp | S_pq | M = pad(x) @ W_g
    const LinPart p_parts[3] = {{w_qp_dst, cin, 0, H}, {w_pq_src, cin, 0, H}, {w_g, H, 1, H}};
    if (enc_gemm_fused(e, st, NP, cin, Xp, ZD, p_parts, 3, Sp, 3 * H)) return 1;
    rowdot_kernel<<<(NQ + 7) / 8, 256, 0, st>>>(Sq, 2 * H, NQ, H, a_qp_src, as_q);        // a_s of q->p
    rowdot_kernel<<<(NQ + 7) / 8, 256, 0, st>>>(Sq + H, 2 * H, NQ, H, a_pq_dst, ad_q);    // a_d of p->q
    rowdot_kernel<<<(NP + 7) / 8, 256, 0, st>>>(Sp, 3 * H, NP, H, a_qp_dst, ad_p);        // a_d of q->p
    rowdot_kernel<<<(NP + 7) / 8, 256, 0, st>>>(Sp + H, 3 * H, NP, H, a_pq_src, as_p);    // a_s of p->q
    // GAT q->p into Gp (pre-activation: the product also receives the GatedGraphConv branch)
    gat_aggregate_kernel<<<NP, 256, 0, st>>>(qp.rowptr, qp.col, Sq, 2 * H, as_q, ad_p, b_qp, H, Gp, H, 0);
    // GAT p->q straight into the next feature block of the queries, relu fused (model/gnn.py:72)
    gat_aggregate_kernel<<<NQ, 256, 0, st>>>(pq.rowptr, pq.col, Sp + H, 3 * H, as_p, ad_q, b_pq, H, Zq + off_next, ZD, 1);
    // GatedGraphConv: aggregate, GRU
    segsum_rows_kernel<<<NP, 256, 0, st>>>(pp.rowptr, pp.col, Sp + 2 * H, 3 * H, H, Agg, H);
    if (enc_gemm(e, st, NP, 3 * H, H, Agg, H, w_ih, H, 0, gi, 3 * H)) return 1;
    if (enc_gemm(e, st, NP, 3 * H, cin, Xp, ZD, w_hh, H, 0, gh, 3 * H)) return 1;  // pad(x) @ W_hh^T = x @ W_hh[:, :cin]^T
    gru_relu_kernel<<<(unsigned)(((int64_t)NP * H + 255) / 256), 256, 0, st>>>(gi, gh, b_ih, b_hh, Xp, ZD, cin, Gp, NP, H,
                                                                               Zp + off_next, ZD);
  }

  if (!run_pool) {
    SSS_CUDA_OK(cudaGetLastError());
    return ws_end(e, st);
  }
  // ---- PositionalAttentionPooling (model/gnn.py:193-217)
  e->a_split.clear();
  const float *wq = P(e, "pooling.query_lin.weight", (int64_t)LIN * ZD), *bq = P(e, "pooling.query_lin.bias", LIN),
              *wp = P(e, "pooling.product_lin.weight", (int64_t)LIN * ZD), *bp = P(e, "pooling.product_lin.bias", LIN),
              *pe = P(e, "pooling.positional_emb.weight", (int64_t)MSL * MSL),
              *wn = P(e, "pooling.node_emb_lin.weight", (int64_t)OUT * OUT), *bn = P(e, "pooling.node_emb_lin.bias", OUT),
              *wc = P(e, "pooling.coarse_rep_lin.weight", (int64_t)OUT * OUT), *wa = P(e, "pooling.att_lin.weight", OUT);
  if (!wq || !bq || !wp || !bp || !pe || !wn || !bn || !wc || !wa) return 1;
  if (enc_gemm(e, st, NQ, LIN, ZD, Zq, ZD, wq, ZD, 0, uq_lin, LIN)) return 1;
  if (enc_gemm(e, st, NP, LIN, ZD, Zp, ZD, wp, ZD, 0, up_lin, LIN)) return 1;
  cnt_to_int_kernel<<<(NP + 255) / 256, 256, 0, st>>>(bt->product_cnt, NP, cnt_i);
  csr_scan_kernel<<<1, 1024, 0, st>>>(cnt_i, NP, cnt_pre, cursor_tmp);
  expand_map_kernel<<<(NP + 127) / 128, 128, 0, st>>>(cnt_pre, NP, occ_prod);
  pool_nodes_kernel<<<NT, 256, 0, st>>>(up_lin, uq_lin, bp, bq, pe, occ_prod, bt->product_pos, bt->query_pos,
                                        bt->product_batch, bt->query_batch, NE, NQ, LIN, MSL, U, node_graph);
  SSS_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(int) * (size_t)B * 4, st));
  graph_ranges_kernel<<<(NT + 255) / 256, 256, 0, st>>>(node_graph, NE, NT, ranges);
  graph_mean_kernel<<<B, 256, 0, st>>>(U, OUT, ranges, nullptr, coarse);
  if (enc_gemm(e, st, NT, OUT, OUT, U, OUT, wn, OUT, 0, Aatt, OUT)) return 1;
  if (enc_gemm(e, st, B, OUT, OUT, coarse, OUT, wc, OUT, 0, Bc, OUT)) return 1;
  pool_att_kernel<<<NT, 256, 0, st>>>(Aatt, bn, Bc, node_graph, wa, OUT, att);
  graph_mean_kernel<<<B, 256, 0, st>>>(U, OUT, ranges, att, out);
  SSS_CUDA_OK(cudaGetLastError());
  return ws_end(e, st);
}

extern "C" int sss_binarize_head(const float* x, const float* W, const float* b, int64_t n, int in_dim, int out_dim,
                                 float* out, int device, void* stream) {
  SSS_REQUIRE(x && W && b && out, "sss_binarize_head: NULL buffer");
  SSS_REQUIRE(n >= 0 && in_dim >= 1 && out_dim >= 1, "sss_binarize_head: bad shape");
  if (n == 0) return 0;
  SSS_REQUIRE(g_cublas.load(), "cuBLAS (libcublas.so.12) could not be loaded");
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  static cublasHandle_t h = nullptr;
  static int h_dev = -1;
  int rc = 0;
  if (!h || h_dev != device) {
    rc = g_cublas.create(&h);
    if (rc == 0) rc = g_cublas.set_math(h, kPedanticMath);
    h_dev = device;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (rc == 0) rc = g_cublas.set_stream(h, st);
  int r2 = rc == 0 ? gemm_nt(h, (int)n, out_dim, in_dim, x, in_dim, W, in_dim, out, out_dim) : 1;
  if (r2 == 0) sign_bias_kernel<<<(unsigned)((n * out_dim + 255) / 256), 256, 0, st>>>(out, b, n, out_dim);
  cudaSetDevice(prev);
  if (rc != 0) set_error("cuBLAS setup failed in sss_binarize_head");
  return (rc != 0 || r2 != 0) ? 1 : 0;
}

