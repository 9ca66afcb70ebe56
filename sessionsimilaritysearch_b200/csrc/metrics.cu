// metrics.cu — the reference's evaluation metrics on the retrieved ids, batched over I[nq, K]:
//   get_score / get_ave_score           fine_tune_ours.py:42-97, 883-897; test_amazon_filterd.py:669-673
// The reference calls get_score nq x K times from Python.  Here:
//   * all_jaccard / cur_jaccard / all_product_type_score are integer set / count work on ragged id lists: one CUDA
//     thread per (query, neighbour) pair over CSR arrays (sss_pair_scores);
//   * all_query_score / all_product_title_score are Levenshtein.seqratio over string lists — nested dynamic
//     programming over code points, "not a GPU problem" (SURVEY 8f): native host code on all hardware threads
//     (sss_seqratio_pairs), restating python-Levenshtein's lev_edit_seq_distance [recalled: the package is not
//     installable here, oracle/metrics_oracle.py restates the same published algorithm].
// Arithmetic follows the reference's numpy float64 path exactly (counts / ||counts||, products summed in numpy's
// pairwise order), then the float32 store into `gt` (fine_tune_ours.py:883-888).
#include <math.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/sss_b200.h"
#include "common.cuh"

namespace sss {

constexpr int PS_MAX_TYPES = 64;

// numpy's float64 add.reduce over a contiguous array (pairwise_sum, n < 128): checked against numpy 2.3 for n < 60
__device__ double np_sum(const double* a, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r += a[i];
    return r;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] += a[i + j];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i];
  return res;
}

// kind 0: Jaccard of two id SETS (lists hold distinct ids): |A & B| / |A | B|, 0 for an empty union
// kind 1: cosine of two type-count vectors built in the reference's type_to_id order (a's types by first appearance,
//         then b's new ones)
__global__ void pair_scores_kernel(int kind, const int64_t* __restrict__ a_off, const int64_t* __restrict__ a_vals,
                                   int64_t nq, const int64_t* __restrict__ b_off, const int64_t* __restrict__ b_vals,
                                   int64_t n_b, const int64_t* __restrict__ I, int k, float* __restrict__ out,
                                   int* __restrict__ err) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * k) return;
  const int64_t q = t / k;
  const int64_t nb_id = I[t];
  if (nb_id < 0 || nb_id >= n_b) {  // padding id (fewer than K results): the reference would index from the end; here 0
    out[t] = 0.0f;
    return;
  }
  const int64_t* A = a_vals + a_off[q];
  const int na = (int)(a_off[q + 1] - a_off[q]);
  const int64_t* B = b_vals + b_off[nb_id];
  const int nb = (int)(b_off[nb_id + 1] - b_off[nb_id]);
  if (kind == 0) {
    int inter = 0;
    for (int i = 0; i < na; ++i) {
      const int64_t v = A[i];
      for (int j = 0; j < nb; ++j) inter += B[j] == v ? 1 : 0;
    }
    const int uni = na + nb - inter;
    out[t] = uni == 0 ? 0.0f : (float)((double)inter / (double)uni);
    return;
  }
  int64_t ids[PS_MAX_TYPES];
  double av[PS_MAX_TYPES], bv[PS_MAX_TYPES];
  int n = 0;
  bool over = false;
  auto slot = [&](int64_t v) {
    for (int i = 0; i < n; ++i)
      if (ids[i] == v) return i;
    if (n == PS_MAX_TYPES) {
      over = true;
      return 0;
    }
    ids[n] = v;
    av[n] = 0.0;
    bv[n] = 0.0;
    return n++;
  };
  for (int i = 0; i < na; ++i) av[slot(A[i])] += 1.0;
  for (int j = 0; j < nb; ++j) bv[slot(B[j])] += 1.0;
  if (over) {
    atomicOr(err, 1);
    out[t] = 0.0f;
    return;
  }
  if (na > 0) {  // a_vec / np.linalg.norm(a_vec): the squared counts are small integers, so the norm is order free
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += av[i] * av[i];
    const double nrm = sqrt(s);
    for (int i = 0; i < n; ++i) av[i] = av[i] / nrm;
  }
  if (nb > 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += bv[i] * bv[i];
    const double nrm = sqrt(s);
    for (int i = 0; i < n; ++i) bv[i] = bv[i] / nrm;
  }
  for (int i = 0; i < n; ++i) av[i] = av[i] * bv[i];
  out[t] = (float)np_sum(av, n);
}

// ---- masked mean over tokens (model/NodeEmbedding.py:113) -----------------------------------------------------
// out[n, h] = sum_t tok[n, t, h] * mask[n, t] / sum_t mask[n, t]; one block per row, threads over h (coalesced),
// t ascending.  HBM-bound: L * H * 4 bytes read per row.
__global__ void masked_mean_kernel(const float* __restrict__ tok, const int64_t* __restrict__ mask, int L, int H,
                                   float* __restrict__ out) {
  const int64_t n = blockIdx.x;
  extern __shared__ float mm_mask[];  // [L]
  __shared__ float s_den;
  for (int t = threadIdx.x; t < L; t += blockDim.x) mm_mask[t] = (float)mask[n * L + t];
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t c = 0;
    for (int t = 0; t < L; ++t) c += mask[n * L + t];
    s_den = (float)c;
  }
  __syncthreads();
  const float* base = tok + n * (int64_t)L * H;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float acc = 0.0f;
    for (int t = 0; t < L; ++t) acc += base[(int64_t)t * H + h] * mm_mask[t];
    out[n * (int64_t)H + h] = acc / s_den;
  }
}

// ---- in-batch cosine matrix -----------------------------------------------------------------------------------
// inverse norms 1 / max(||x||, 1e-12) (F.normalize), one warp per row
__global__ void inv_norm_kernel(const float* __restrict__ x, int64_t n, int d, float* __restrict__ inv) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float ss = 0.0f;
  for (int j = lane; j < d; j += 32) ss = fmaf(x[row * d + j], x[row * d + j], ss);
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) inv[row] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
}
// out[i, j] = sum_k (a[i, k] * ia[i]) * (b[j, k] * ib[j]), k ascending; 16 x 16 outputs per block through shared tiles
__global__ void __launch_bounds__(256) cosine_matrix_kernel(const float* __restrict__ a, const float* __restrict__ ia,
                                                            int64_t na, const float* __restrict__ b,
                                                            const float* __restrict__ ib, int64_t nb, int d,
                                                            float* __restrict__ out) {
  __shared__ float As[16][33], Bs[16][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * 16, j0 = (int64_t)blockIdx.x * 16;
  float acc = 0.0f;
  for (int k0 = 0; k0 < d; k0 += 32) {
    for (int e = threadIdx.x; e < 16 * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      As[r][c] = (i0 + r < na && k0 + c < d) ? a[(i0 + r) * d + k0 + c] * ia[i0 + r] : 0.0f;
      Bs[r][c] = (j0 + r < nb && k0 + c < d) ? b[(j0 + r) * d + k0 + c] * ib[j0 + r] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 32; ++c) acc = fmaf(As[ty][c], Bs[tx][c], acc);
    __syncthreads();
  }
  if (i0 + ty < na && j0 + tx < nb) out[(i0 + ty) * nb + j0 + tx] = acc;
}

// ---- seqratio (host) --------------------------------------------------------------------------------------
// lev_edit_distance(s1, s2, xcost = 1): insert / delete 1, substitute 2
static size_t edit_distance_x1(const uint32_t* a, size_t la, const uint32_t* b, size_t lb, std::vector<size_t>& row) {
  while (la > 0 && lb > 0 && a[0] == b[0]) { ++a; ++b; --la; --lb; }
  while (la > 0 && lb > 0 && a[la - 1] == b[lb - 1]) { --la; --lb; }
  if (la == 0) return lb;
  if (lb == 0) return la;
  if (la > lb) {
    std::swap(a, b);
    std::swap(la, lb);
  }
  row.resize(la + 1);
  for (size_t i = 0; i <= la; ++i) row[i] = i;
  for (size_t j = 1; j <= lb; ++j) {
    size_t prev = row[0];
    row[0] = j;
    const uint32_t cb = b[j - 1];
    for (size_t i = 1; i <= la; ++i) {
      const size_t cur = row[i];
      size_t v = prev + (a[i - 1] == cb ? 0 : 2);
      v = std::min(v, std::min(row[i] + 1, row[i - 1] + 1));
      row[i] = v;
      prev = cur;
    }
  }
  return row[la];
}

struct StrRef {
  const uint32_t* p;
  size_t n;
  bool operator==(const StrRef& o) const { return n == o.n && std::equal(p, p + n, o.p); }
};

// lev_edit_seq_distance [recalled], including its quirk: a cell whose two strings are both empty does not advance the
// inner string pointer
static double edit_seq_distance(std::vector<StrRef> s1, std::vector<StrRef> s2, std::vector<size_t>& scratch,
                                std::vector<double>& row) {
  size_t b1 = 0, e1 = s1.size(), b2 = 0, e2 = s2.size();
  while (b1 < e1 && b2 < e2 && s1[b1] == s2[b2]) { ++b1; ++b2; }
  while (b1 < e1 && b2 < e2 && s1[e1 - 1] == s2[e2 - 1]) { --e1; --e2; }
  if (b1 == e1) return (double)(e2 - b2);
  if (b2 == e2) return (double)(e1 - b1);
  const StrRef *x = s1.data() + b1, *y = s2.data() + b2;
  size_t n1 = e1 - b1, n2 = e2 - b2;
  if (n1 > n2) {
    std::swap(x, y);
    std::swap(n1, n2);
  }
  ++n1;
  ++n2;
  row.resize(n2);
  for (size_t i = 0; i < n2; ++i) row[i] = (double)i;
  for (size_t i = 1; i < n1; ++i) {
    const StrRef& a = x[i - 1];
    double D = (double)i - 1.0;
    double v = (double)i;
    size_t j2 = 0;
    for (size_t p = 1; p < n2; ++p) {
      const StrRef& b = y[j2];
      const size_t l = a.n + b.n;
      double q;
      if (l == 0) {
        q = D;
      } else {
        const size_t d = edit_distance_x1(a.p, a.n, b.p, b.n, scratch);
        ++j2;
        q = D + 2.0 / (double)l * (double)d;
      }
      v += 1.0;
      if (v > q) v = q;
      D = row[p];
      if (v > D + 1.0) v = D + 1.0;
      row[p] = v;
    }
  }
  return row[n2 - 1];
}

static void gather_seq(const sss_string_seqs_t* s, int64_t i, std::vector<StrRef>& out) {
  out.clear();
  for (int64_t t = s->seq_off[i]; t < s->seq_off[i + 1]; ++t)
    out.push_back(StrRef{s->chars + s->str_off[t], (size_t)(s->str_off[t + 1] - s->str_off[t])});
}

}  // namespace sss

using namespace sss;

extern "C" int sss_pair_scores(int kind, const int64_t* a_off, const int64_t* a_vals, int64_t nq, const int64_t* b_off,
                               const int64_t* b_vals, int64_t n_b, const int64_t* I, int k, float* out, int device,
                               void* stream) {
  SSS_REQUIRE(kind == SSS_SCORE_JACCARD || kind == SSS_SCORE_TYPE_COSINE, "sss_pair_scores: unknown kind");
  SSS_REQUIRE(nq >= 0 && k >= 1 && n_b >= 0, "sss_pair_scores: bad shape");
  if (nq == 0) return 0;
  SSS_REQUIRE(a_off && b_off && I && out, "sss_pair_scores: NULL buffer");  // (a_vals / b_vals may be NULL when every list is empty)
  int prev = 0;
  SSS_CUDA_OK(cudaGetDevice(&prev));
  SSS_CUDA_OK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  int* flag = nullptr;
  int host_flag = 0;
  cudaError_t err = cudaMalloc((void**)&flag, sizeof(int));
  if (err == cudaSuccess) err = cudaMemsetAsync(flag, 0, sizeof(int), st);
  if (err == cudaSuccess) {
    const int64_t total = nq * k;
    pair_scores_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(kind, a_off, a_vals, nq, b_off, b_vals, n_b, I, k,
                                                                        out, flag);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (flag) cudaFree(flag);
  cudaSetDevice(prev);
  if (err != cudaSuccess) {
    set_error(std::string("sss_pair_scores: ") + cudaGetErrorString(err));
    return 1;
  }
  SSS_REQUIRE(host_flag == 0, "sss_pair_scores: more than 64 distinct product types in one pair of sessions");
  return 0;
}

extern "C" int sss_masked_mean(const float* tok, const int64_t* mask, int64_t n, int L, int H, float* out, int device,
                               void* stream) {
  SSS_REQUIRE(n >= 0 && L >= 1 && H >= 1, "sss_masked_mean: bad shape");
  if (n == 0) return 0;
  SSS_REQUIRE(tok && mask && out, "sss_masked_mean: NULL buffer");
  int prev = 0;
  SSS_CUDA_OK(cudaGetDevice(&prev));
  SSS_CUDA_OK(cudaSetDevice(device));
  masked_mean_kernel<<<(unsigned)n, 256, (size_t)L * sizeof(float), (cudaStream_t)stream>>>(tok, mask, L, H, out);
  cudaError_t err = cudaGetLastError();
  cudaSetDevice(prev);
  SSS_REQUIRE(err == cudaSuccess, std::string("sss_masked_mean: ") + cudaGetErrorString(err));
  return 0;
}

extern "C" int sss_cosine_matrix(const float* a, int64_t na, const float* b, int64_t nb, int d, float* out, int device,
                                 void* stream) {
  SSS_REQUIRE(na >= 0 && nb >= 0 && d >= 1, "sss_cosine_matrix: bad shape");
  if (na == 0 || nb == 0) return 0;
  SSS_REQUIRE(a && b && out, "sss_cosine_matrix: NULL buffer");
  int prev = 0;
  SSS_CUDA_OK(cudaGetDevice(&prev));
  SSS_CUDA_OK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  float* inv = nullptr;
  cudaError_t err = cudaMallocAsync((void**)&inv, (size_t)(na + nb) * sizeof(float), st);
  if (err == cudaSuccess) {
    inv_norm_kernel<<<(unsigned)((na + 7) / 8), 256, 0, st>>>(a, na, d, inv);
    inv_norm_kernel<<<(unsigned)((nb + 7) / 8), 256, 0, st>>>(b, nb, d, inv + na);
    dim3 grid((unsigned)((nb + 15) / 16), (unsigned)((na + 15) / 16));
    cosine_matrix_kernel<<<grid, 256, 0, st>>>(a, inv, na, b, inv + na, nb, d, out);
    err = cudaGetLastError();
    cudaFreeAsync(inv, st);
  }
  cudaSetDevice(prev);
  SSS_REQUIRE(err == cudaSuccess, std::string("sss_cosine_matrix: ") + cudaGetErrorString(err));
  return 0;
}

extern "C" int sss_seqratio_pairs(const sss_string_seqs_t* a, const sss_string_seqs_t* b, const int64_t* I, int64_t nq,
                                  int k, int zero_if_empty, float* out, int n_threads) {
  SSS_REQUIRE(a && b && I && out, "sss_seqratio_pairs: NULL argument");
  SSS_REQUIRE(nq >= 0 && k >= 1 && a->n_seqs >= nq, "sss_seqratio_pairs: bad shape");
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, nq));
  auto work = [&](int tid) {
    std::vector<StrRef> sa, sb;
    std::vector<size_t> scratch;
    std::vector<double> row;
    for (int64_t q = tid; q < nq; q += n_threads) {
      gather_seq(a, q, sa);
      for (int j = 0; j < k; ++j) {
        const int64_t id = I[q * k + j];
        float v = 0.0f;
        if (id >= 0 && id < b->n_seqs) {
          gather_seq(b, id, sb);
          if (!(zero_if_empty && (sa.empty() || sb.empty()))) {
            const size_t lensum = sa.size() + sb.size();
            const double r = lensum == 0 ? 1.0 : ((double)lensum - edit_seq_distance(sa, sb, scratch, row)) / (double)lensum;
            v = (float)r;
          }
        }
        out[q * k + j] = v;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  return 0;
}
