// encoder.cu — the session encoder forward after the text embedder:
//   UnifyPoolingGraphLevelEncoder.forward   model/model.py:279-351 (use_id_embedding=False, eval)
//   HeteroGGNN.forward                      model/gnn.py:64-81
//   GATConv / GatedGraphConv / HeteroConv   torch_geometric 2.0.4 [recalled], SURVEY.md Appendix A
//   PositionalAttentionPooling.forward      model/gnn.py:193-217
//   BinarizeHead.forward (eval, mlp=None)   model/model.py:117-138
//
// 16 launches per forward (was ~90), no library GEMM:
//   ingest          x -> node-embedding block 0 as fp32 and as the hi / lo bf16 GEMM operand, NaN flag
//   graph_prep      4 blocks: destination-sorted CSR of the three edge types (GATConv's bipartite self-loop quirk
//                   included) and the pooling's occurrence prefix / graph ranges
//   per layer (x3): ONE dual GEMM launch (query-side [S_qp | T_pq] and product-side [T_qp | S_pq | M | gh] linears,
//                   attention logits as epilogue partials), ONE message-passing launch (both GAT directions: edge
//                   softmax computed once per destination, float4 weighted segment sums; GatedGraphConv segment sum
//                   written straight as the next GEMM's operand), ONE GRU GEMM (gates + HeteroConv sum + relu in
//                   the epilogue, next layer's features written as fp32 and hi / lo)
//   pooling:        ONE dual GEMM (query_lin / product_lin with bias, repeat_interleave, positional concat, tanh in
//                   the epilogue), graph mean, coarse_rep_lin GEMM, node_emb_lin GEMM with the gated attention logit
//                   as epilogue partials, weighted graph mean
// All dense linears run on this library's split-bf16 tcgen05 GEMM (gemm_bf16x3_sm100.cu): 2.7e-5 of the output scale
// from a float64 forward.  Segment reductions and the partial sums walk their terms in a fixed order, so a forward
// is deterministic run to run.
#include <math.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "../../include/sss_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace sss {

__device__ __forceinline__ void put_hilo(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ---- ingest: src[n, w] (row pitch ld_src) -> Z[:, 0:w) fp32 (row pitch ldz; skipped when src == Z) and hi / lo ----
__global__ void ingest_kernel(const float* __restrict__ srcq, int nq, const float* __restrict__ srcp, int np, int w,
                              int ld_src, float* __restrict__ zq, float* __restrict__ zp, int ldz,
                              __nv_bfloat16* __restrict__ qhi, __nv_bfloat16* __restrict__ qlo,
                              __nv_bfloat16* __restrict__ phi, __nv_bfloat16* __restrict__ plo, int ldh,
                              int32_t* __restrict__ nonfinite) {
  const int r = blockIdx.x;
  const bool is_q = r < nq;
  const int row = is_q ? r : r - nq;
  const float* src = (is_q ? srcq : srcp) + (size_t)row * ld_src;
  float* z = (is_q ? zq : zp) + (size_t)row * ldz;
  __nv_bfloat16* hi = is_q ? qhi : phi;
  __nv_bfloat16* lo = is_q ? qlo : plo;
  bool bad = false;
  for (int c = threadIdx.x; c < w; c += blockDim.x) {
    const float v = src[c];
    bad |= isnan(v);
    if (z != src) z[c] = v;
    put_hilo(hi, lo, (size_t)row * ldh + c, v);
  }
  if (bad && nonfinite) *nonfinite = 1;
}

// ---- graph_prep: one block per job -----------------------------------------------------------------------
// jobs 0..2: CSR by destination of one edge type.  GATConv(add_self_loops=True) on a bipartite edge set: drop edges
// with src == dst (batch-global indices), then append (i, i) for i < n_loop = min(N_src, N_dst); GatedGraphConv:
// n_loop = 0, nothing dropped.  Lists are sorted by source, which fixes the summation order.
// job 3: the pooling's structure: prefix of the occurrence counts, row -> graph map, per-graph row ranges.
struct CsrJob {
  const int64_t *src, *dst;
  int64_t E;
  int n_dst, drop_self, n_loop;
  int *deg, *rowptr, *col;  // deg: [n_dst + 1] scratch (doubles as the fill cursor)
};
struct PrepArgs {
  CsrJob csr[3];
  int run_csr;
  const int64_t *cnt, *product_batch, *query_batch;
  int n_p, n_q, n_e, n_graphs;
  int *prefix, *node_graph, *ranges, *cnt_i;  // [n_p + 1], [n_e + n_q], [n_graphs * 4], [n_p + 1] scratch
};

// exclusive scan of v[0, n) into out[0, n], whole block (1024 threads)
__device__ void block_exclusive_scan(const int* __restrict__ v, int n, int* __restrict__ out) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const int x = i < n ? v[i] : 0;
    int inc = x;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;  // inclusive
    }
    __syncthreads();
    const int before = carry_s + (warp > 0 ? warp_tot[warp - 1] : 0) + inc - x;
    if (i < n) out[i] = before;
    __syncthreads();
    if (tid == 1023) carry_s = before + x;
    __syncthreads();
  }
  if (tid == 0) out[n] = carry_s;
  __syncthreads();
}

__global__ void __launch_bounds__(1024) graph_prep_kernel(PrepArgs a) {
  const int tid = threadIdx.x;
  if (blockIdx.x < 3) {
    if (!a.run_csr) return;
    const CsrJob j = a.csr[blockIdx.x];
    for (int i = tid; i <= j.n_dst; i += 1024) j.deg[i] = 0;
    __syncthreads();
    const int64_t tot = j.E + j.n_loop;
    for (int64_t t = tid; t < tot; t += 1024) {
      if (t < j.E) {
        if (!(j.drop_self && j.src[t] == j.dst[t])) atomicAdd(&j.deg[j.dst[t]], 1);
      } else {
        atomicAdd(&j.deg[t - j.E], 1);
      }
    }
    __syncthreads();
    block_exclusive_scan(j.deg, j.n_dst, j.rowptr);
    for (int i = tid; i < j.n_dst; i += 1024) j.deg[i] = j.rowptr[i];  // fill cursors
    __syncthreads();
    for (int64_t t = tid; t < tot; t += 1024) {
      if (t < j.E) {
        if (!(j.drop_self && j.src[t] == j.dst[t])) j.col[atomicAdd(&j.deg[j.dst[t]], 1)] = (int)j.src[t];
      } else {
        const int i = (int)(t - j.E);
        j.col[atomicAdd(&j.deg[i], 1)] = i;
      }
    }
    __syncthreads();
    for (int i = tid; i < j.n_dst; i += 1024) {  // lists are tiny: insertion sort by source
      const int b = j.rowptr[i], e = j.rowptr[i + 1];
      for (int x = b + 1; x < e; ++x) {
        const int v = j.col[x];
        int y = x - 1;
        while (y >= b && j.col[y] > v) {
          j.col[y + 1] = j.col[y];
          --y;
        }
        j.col[y + 1] = v;
      }
    }
    return;
  }
  // ---- pooling structure
  for (int i = tid; i < a.n_p; i += 1024) a.cnt_i[i] = (int)a.cnt[i];
  for (int i = tid; i < a.n_graphs * 4; i += 1024) a.ranges[i] = 0;
  __syncthreads();
  block_exclusive_scan(a.cnt_i, a.n_p, a.prefix);
  const int n_tot = a.n_e + a.n_q;
  for (int p = tid; p < a.n_p; p += 1024)
    for (int t = a.prefix[p]; t < a.prefix[p + 1] && t < a.n_e; ++t) a.node_graph[t] = (int)a.product_batch[p];
  for (int q = tid; q < a.n_q; q += 1024) a.node_graph[a.n_e + q] = (int)a.query_batch[q];
  __syncthreads();
  // graph -> [first, last) rows among the product occurrences and among the queries (both batch vectors are sorted)
  for (int r = tid; r < n_tot; r += 1024) {
    const int g = a.node_graph[r];
    const int part = r < a.n_e ? 0 : 1;
    const bool first = (r == 0) || (r == a.n_e) || a.node_graph[r - 1] != g;
    const bool last = (r == a.n_e - 1) || (r == n_tot - 1) || a.node_graph[r + 1] != g;
    if (first) a.ranges[g * 4 + part * 2] = r;
    if (last) a.ranges[g * 4 + part * 2 + 1] = r + 1;
  }
}

// ---- message passing of one HeteroConv layer, one block per destination node ---------------------------------
// product p:  Gp[p, :]  = sum_e alpha_e * S_qp[src_e, :] + b_qp        (GAT query -> product, pre-activation)
//             Agg[p, :] = sum_e M[src_e, :]                            (GatedGraphConv, written as hi / lo bf16)
// query q:    Xq_next[q, :] = relu(sum_e alpha_e * S_pq[src_e, :] + b_pq)   (GAT product -> query; fp32 + hi / lo)
// alpha = softmax over the incoming edges of leaky_relu(a_s[src] + a_d[dst], 0.2) with PyG's denominator
// (sum + 1e-16); a destination without incoming edges gets the bias.  a_s / a_d arrive as per-tile partials of the
// GEMM epilogue and are summed here in tile order.  The edge weights are computed ONCE per destination (warp 0) and
// every thread then accumulates four columns per edge with 16-byte loads.
struct MpArgs {
  int n_p, n_q, H, parts;      // parts = partial sums per attention scalar
  const int *qp_rowptr, *qp_col, *pq_rowptr, *pq_col, *pp_rowptr, *pp_col;
  const float* Sq; int ldsq;   // [n_q, 2 * PP]: S_qp | T_pq
  const float* Sp; int ldsp;   // [n_p, 3 * PP + GH]: T_qp | S_pq | M | gh
  int PP;
  const float *asq, *adq, *adp, *asp;  // partials [n, parts]
  const float *b_qp, *b_pq;
  float* Gp;                   // [n_p, H]
  __nv_bfloat16 *agg_hi, *agg_lo; int ld_agg;  // [n_p_pad, pad64(H)]
  float* zq_next; int ldz;     // Zq + off_next
  __nv_bfloat16 *zq_hi, *zq_lo; int ldh, k0_next;
};
constexpr int kMpEdges = 128;

__device__ __forceinline__ float sum_parts(const float* p, int n) {
  float s = 0.0f;
  for (int i = 0; i < n; ++i) s += p[i];
  return s;
}

// edge softmax of one destination: alpha[e - b] for e in [b, min(e_end, b + kMpEdges)) plus the global max / den
__device__ void edge_softmax(const int* col, int b, int e, const float* as_part, float ad, int parts, float* s_alpha,
                             int* s_src, float* s_stat) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    float mx = -INFINITY;
    for (int k = b + lane; k < e; k += 32) {
      float v = sum_parts(as_part + (size_t)col[k] * parts, parts) + ad;
      v = v > 0.0f ? v : 0.2f * v;
      if (k - b < kMpEdges) {
        s_alpha[k - b] = v;
        s_src[k - b] = col[k];
      }
      mx = fmaxf(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // the denominator is summed in edge order by one lane: the order is part of the result
    float den = 0.0f;
    if (lane == 0) {
      for (int k = b; k < e; ++k) {
        float v;
        if (k - b < kMpEdges) {
          v = s_alpha[k - b];
        } else {
          v = sum_parts(as_part + (size_t)col[k] * parts, parts) + ad;
          v = v > 0.0f ? v : 0.2f * v;
        }
        den += expf(v - mx);
      }
      den += 1e-16f;
      s_stat[0] = mx;
      s_stat[1] = den;
    }
    den = __shfl_sync(0xffffffffu, den, 0);
    for (int k = b + lane; k < e && k - b < kMpEdges; k += 32) s_alpha[k - b] = expf(s_alpha[k - b] - mx) / den;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) message_passing_kernel(MpArgs a) {
  __shared__ float s_alpha[kMpEdges];
  __shared__ int s_src[kMpEdges];
  __shared__ float s_stat[2];
  const int node = blockIdx.x;
  const bool is_p = node < a.n_p;
  const int i = is_p ? node : node - a.n_p;
  const int H = a.H;
  // ---- GAT into this destination
  const int* rowptr = is_p ? a.qp_rowptr : a.pq_rowptr;
  const int* col = is_p ? a.qp_col : a.pq_col;
  const int b = rowptr[i], e = rowptr[i + 1];
  const float ad = sum_parts((is_p ? a.adp : a.adq) + (size_t)i * a.parts, a.parts);
  const float* as_part = is_p ? a.asq : a.asp;
  edge_softmax(col, b, e, as_part, ad, a.parts, s_alpha, s_src, s_stat);
  const float* S = is_p ? a.Sq : a.Sp + a.PP;   // source features: S_qp (queries) / S_pq (products)
  const int lds = is_p ? a.ldsq : a.ldsp;
  const float* bias = is_p ? a.b_qp : a.b_pq;
  const int ne = e - b;
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = c + 4 <= H;
    for (int k = 0; k < ne; ++k) {
      int j;
      float al;
      if (k < kMpEdges) {
        j = s_src[k];
        al = s_alpha[k];
      } else {  // destinations with more than kMpEdges incoming edges: weights recomputed on the fly
        j = col[b + k];
        float v = sum_parts(as_part + (size_t)j * a.parts, a.parts) + ad;
        v = v > 0.0f ? v : 0.2f * v;
        al = expf(v - s_stat[0]) / s_stat[1];
      }
      const float* sr = S + (size_t)j * lds + c;
      if (vec) {
        const float4 x = *reinterpret_cast<const float4*>(sr);
        acc[0] += x.x * al; acc[1] += x.y * al; acc[2] += x.z * al; acc[3] += x.w * al;
      } else {
        for (int t = 0; t < 4 && c + t < H; ++t) acc[t] += sr[t] * al;
      }
    }
    for (int t = 0; t < 4 && c + t < H; ++t) {
      const float v = acc[t] + bias[c + t];
      if (is_p) {
        a.Gp[(size_t)i * H + c + t] = v;
      } else {
        const float o = fmaxf(v, 0.0f);  // relu of the layer output (model/gnn.py:72)
        a.zq_next[(size_t)i * a.ldz + c + t] = o;
        put_hilo(a.zq_hi, a.zq_lo, (size_t)i * a.ldh + a.k0_next + c + t, o);
      }
    }
  }
  if (!is_p) return;
  // ---- GatedGraphConv aggregation into this product: Agg = sum of M over incoming product -> product edges
  const int pb = a.pp_rowptr[i], pe = a.pp_rowptr[i + 1];
  const float* M = a.Sp + 2 * a.PP;
  for (int c = threadIdx.x * 4; c < a.ld_agg; c += blockDim.x * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < H) {
      const bool vec = c + 4 <= H;
      for (int k = pb; k < pe; ++k) {
        const float* sr = M + (size_t)a.pp_col[k] * a.ldsp + c;
        if (vec) {
          const float4 x = *reinterpret_cast<const float4*>(sr);
          acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
        } else {
          for (int t = 0; t < 4 && c + t < H; ++t) acc[t] += sr[t];
        }
      }
    }
    for (int t = 0; t < 4 && c + t < a.ld_agg; ++t)
      put_hilo(a.agg_hi, a.agg_lo, (size_t)i * a.ld_agg + c + t, c + t < H ? acc[t] : 0.0f);  // zero K padding
  }
}

// ---- pooling ---------------------------------------------------------------------------------------------
// mean over the rows of a graph: coarse[g, :] = sum_r U[r, :] / count, written as the hi / lo operand of coarse_rep_lin
__global__ void graph_mean_hilo_kernel(const float* __restrict__ U, int W, const int* __restrict__ ranges,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ldh) {
  const int g = blockIdx.x;
  const int p0 = ranges[g * 4], p1 = ranges[g * 4 + 1], q0 = ranges[g * 4 + 2], q1 = ranges[g * 4 + 3];
  const float cnt = fmaxf((float)((p1 - p0) + (q1 - q0)), 1.0f);
  for (int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4; c < W; c += gridDim.y * blockDim.x * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = c + 4 <= W;
    for (int pass = 0; pass < 2; ++pass) {
      const int r0 = pass ? q0 : p0, r1 = pass ? q1 : p1;
      for (int r = r0; r < r1; ++r) {
        const float* u = U + (size_t)r * W + c;
        if (vec) {
          const float4 x = *reinterpret_cast<const float4*>(u);
          acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
        } else {
          for (int t = 0; t < 4 && c + t < W; ++t) acc[t] += u[t];
        }
      }
    }
    for (int t = 0; t < 4 && c + t < W; ++t) put_hilo(hi, lo, (size_t)g * ldh + c + t, acc[t] / cnt);
  }
}
// out[g, :] = sum_r att[r] * U[r, :] / count, att[r] = the epilogue's per-tile partials summed in tile order
__global__ void graph_weighted_mean_kernel(const float* __restrict__ U, int W, const int* __restrict__ ranges,
                                           const float* __restrict__ att_part, int parts, float* __restrict__ out) {
  extern __shared__ float s_att[];  // attention logits of this graph's rows
  const int g = blockIdx.x;
  const int p0 = ranges[g * 4], p1 = ranges[g * 4 + 1], q0 = ranges[g * 4 + 2], q1 = ranges[g * 4 + 3];
  const int np = p1 - p0, nr = np + (q1 - q0);
  const float cnt = fmaxf((float)nr, 1.0f);
  for (int base = 0; base < nr; base += 1024) {  // graphs of more than 1024 expanded nodes: chunks
    const int m = min(1024, nr - base);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const int k = base + i;
      const int r = k < np ? p0 + k : q0 + (k - np);
      s_att[i] = sum_parts(att_part + (size_t)r * parts, parts);
    }
    __syncthreads();
    for (int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4; c < W; c += gridDim.y * blockDim.x * 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const bool vec = c + 4 <= W;
      if (base > 0)
        for (int t = 0; t < 4 && c + t < W; ++t) acc[t] = out[(size_t)g * W + c + t];
      for (int i = 0; i < m; ++i) {
        const int k = base + i;
        const int r = k < np ? p0 + k : q0 + (k - np);
        const float* u = U + (size_t)r * W + c;
        const float w = s_att[i];
        if (vec) {
          const float4 x = *reinterpret_cast<const float4*>(u);
          acc[0] += x.x * w; acc[1] += x.y * w; acc[2] += x.z * w; acc[3] += x.w * w;
        } else {
          for (int t = 0; t < 4 && c + t < W; ++t) acc[t] += u[t] * w;
        }
      }
      const bool last = base + m >= nr;
      for (int t = 0; t < 4 && c + t < W; ++t) out[(size_t)g * W + c + t] = last ? acc[t] / cnt : acc[t];
    }
  }
  if (nr == 0 && blockIdx.y == 0)
    for (int c = threadIdx.x; c < W; c += blockDim.x) out[(size_t)g * W + c] = 0.0f;
}

}  // namespace sss

using namespace sss;

namespace {
struct SplitW {  // a weight operand: hi / lo bf16 [rows_pad, k_pad]
  void *hi = nullptr, *lo = nullptr;
  int rows_pad = 0, k_pad = 0;
};
inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }
}  // namespace

struct sss_encoder {
  int device = 0;
  sss_encoder_shape_t sh{};
  std::map<std::string, float*> params;
  std::map<std::string, int64_t> numel;
  // split-bf16 weight operands, built on the first forward after the parameters were (re)set
  bool weights_ready = false;
  std::vector<SplitW> w_q, w_p, w_ih;  // per layer: query-side cat, product-side cat, permuted GRU input weights
  SplitW w_poolq, w_poolp, w_node, w_coarse;
  std::vector<void*> w_allocs;
  int* gemm_flag = nullptr;  // watchdog code of the tensor-core GEMM
  // Workspace arena: slabs are bump-allocated per forward call and kept across calls (cudaMalloc/cudaFree per
  // buffer cost more than the whole forward).  A call that needed more than one slab is followed by one
  // consolidation at the start of the next call; `done` orders reuse across streams.
  std::vector<std::pair<char*, size_t>> slabs;
  size_t used = 0;        // bytes taken from the last slab
  size_t requested = 0;   // bytes requested by the current / last call
  cudaEvent_t done = nullptr;
  bool done_recorded = false;
  int64_t launches = 0;   // kernels of the last forward
  int64_t host_ns = 0;    // host time the last forward spent enqueueing them
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
void drop_weights(sss_encoder* e) {
  for (void* p : e->w_allocs) cudaFree(p);
  e->w_allocs.clear();
  e->w_q.clear();
  e->w_p.clear();
  e->w_ih.clear();
  e->weights_ready = false;
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

// one part of a concatenated weight: n source rows (x @ W.T: rows of W; transposed: columns of W), written as
// rows_out rows (zero beyond n)
struct WPart {
  const float* W;
  int ld, transposed, n, rows_out;
};
int build_weight(sss_encoder* e, const WPart* parts, int n_parts, int K, SplitW* out, cudaStream_t st,
                 const int* row_map = nullptr, int rows_mapped = 0) {
  int rows = 0;
  for (int i = 0; i < n_parts; ++i) rows += parts[i].rows_out;
  if (row_map) rows = rows_mapped;
  out->rows_pad = pad_to(rows, 128);
  out->k_pad = pad_to(K, 64);
  const size_t bytes = (size_t)out->rows_pad * out->k_pad * 2;
  SSS_CUDA_OK(cudaMalloc(&out->hi, bytes));
  e->w_allocs.push_back(out->hi);
  SSS_CUDA_OK(cudaMalloc(&out->lo, bytes));
  e->w_allocs.push_back(out->lo);
  SSS_CUDA_OK(cudaMemsetAsync(out->hi, 0, bytes, st));
  SSS_CUDA_OK(cudaMemsetAsync(out->lo, 0, bytes, st));
  if (row_map) {
    return launch_split_bf16(parts[0].W, parts[0].n, K, parts[0].ld, parts[0].transposed, out->hi, out->lo, rows_mapped,
                             out->k_pad, st, row_map);
  }
  int off = 0;
  for (int i = 0; i < n_parts; ++i) {
    if (launch_split_bf16(parts[i].W, parts[i].n, K, parts[i].ld, parts[i].transposed,
                          (uint16_t*)out->hi + (size_t)off * out->k_pad, (uint16_t*)out->lo + (size_t)off * out->k_pad,
                          parts[i].rows_out, out->k_pad, st))
      return 1;
    off += parts[i].rows_out;
  }
  return 0;
}

int prepare_weights(sss_encoder* e, cudaStream_t st) {
  if (e->weights_ready) return 0;
  drop_weights(e);
  const int IN = e->sh.in_dim, H = e->sh.hidden, L = e->sh.n_layers, OUT = e->sh.out_dim, MSL = e->sh.max_seq_len;
  const int LIN = OUT - MSL, ZD = IN + L * H;
  const int PP = pad_to(H, 128), GH = pad_to(3 * H, 128);
  e->w_q.resize(L);
  e->w_p.resize(L);
  e->w_ih.resize(L);
  // GRU input weights permuted so that an N tile of 96 rows holds the r, z, n rows of the same 32 hidden units
  const int n_ut = (H + 31) / 32;
  std::vector<int> map((size_t)n_ut * 96, -1);
  for (int t = 0; t < n_ut; ++t)
    for (int gate = 0; gate < 3; ++gate)
      for (int i = 0; i < 32; ++i)
        if (t * 32 + i < H) map[(size_t)t * 96 + gate * 32 + i] = gate * H + t * 32 + i;
  int* d_map = nullptr;
  SSS_CUDA_OK(cudaMalloc((void**)&d_map, map.size() * sizeof(int)));
  e->w_allocs.push_back(d_map);
  SSS_CUDA_OK(cudaMemcpyAsync(d_map, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  for (int l = 0; l < L; ++l) {
    const int cin = l == 0 ? IN : H;
    const std::string pre = "gnn.convs." + std::to_string(l) + ".convs.";
    const std::string eqp = pre + "query__clicks__product.", epq = pre + "product__clicked by__query.",
                      epp = pre + "product__to__product.";
    const float *w_qp_src = P(e, eqp + "lin_src.weight", (int64_t)H * cin), *w_qp_dst = P(e, eqp + "lin_dst.weight", (int64_t)H * cin),
                *w_pq_src = P(e, epq + "lin_src.weight", (int64_t)H * cin), *w_pq_dst = P(e, epq + "lin_dst.weight", (int64_t)H * cin),
                *w_g = P(e, epp + "weight", (int64_t)H * H), *w_ih = P(e, epp + "rnn.weight_ih", (int64_t)3 * H * H),
                *w_hh = P(e, epp + "rnn.weight_hh", (int64_t)3 * H * H);
    if (!w_qp_src || !w_qp_dst || !w_pq_src || !w_pq_dst || !w_g || !w_ih || !w_hh) return 1;
    // query side: S_qp (source of q->p) | T_pq (destination of p->q), each padded to PP rows
    const WPart qparts[2] = {{w_qp_src, cin, 0, H, PP}, {w_pq_dst, cin, 0, H, PP}};
    if (build_weight(e, qparts, 2, cin, &e->w_q[l], st)) return 1;
    // product side: T_qp | S_pq | M = pad(x) @ W_g | gh = pad(x) @ W_hh^T = x @ W_hh[:, :cin]^T
    const WPart pparts[4] = {{w_qp_dst, cin, 0, H, PP}, {w_pq_src, cin, 0, H, PP}, {w_g, H, 1, H, PP}, {w_hh, H, 0, 3 * H, GH}};
    if (build_weight(e, pparts, 4, cin, &e->w_p[l], st)) return 1;
    const WPart ih = {w_ih, H, 0, 3 * H, 0};
    if (build_weight(e, &ih, 1, H, &e->w_ih[l], st, d_map, n_ut * 96)) return 1;
  }
  const float *wq = P(e, "pooling.query_lin.weight", (int64_t)LIN * ZD), *wp = P(e, "pooling.product_lin.weight", (int64_t)LIN * ZD),
              *wn = P(e, "pooling.node_emb_lin.weight", (int64_t)OUT * OUT), *wc = P(e, "pooling.coarse_rep_lin.weight", (int64_t)OUT * OUT);
  if (!wq || !wp || !wn || !wc) return 1;
  const WPart pq_ = {wq, ZD, 0, LIN, pad_to(OUT, 128)}, pp_ = {wp, ZD, 0, LIN, pad_to(OUT, 128)},
              pn_ = {wn, OUT, 0, OUT, pad_to(OUT, 128)}, pc_ = {wc, OUT, 0, OUT, pad_to(OUT, 128)};
  if (build_weight(e, &pq_, 1, ZD, &e->w_poolq, st) || build_weight(e, &pp_, 1, ZD, &e->w_poolp, st) ||
      build_weight(e, &pn_, 1, OUT, &e->w_node, st) || build_weight(e, &pc_, 1, OUT, &e->w_coarse, st))
    return 1;
  if (!e->gemm_flag) {
    SSS_CUDA_OK(cudaMalloc((void**)&e->gemm_flag, sizeof(int)));
    e->w_allocs.push_back(e->gemm_flag);
    SSS_CUDA_OK(cudaMemsetAsync(e->gemm_flag, 0, sizeof(int), st));
  }
  e->weights_ready = true;
  return 0;
}

// common fields of a GEMM problem: A [M rows, row pitch a_ld] read from column a_k0 over K columns; B = a SplitW
GemmProblem make_problem(const void* a_hi, const void* a_lo, int M, int a_ld, int a_k0, int K, const SplitW& w, int N,
                         int bn, int epi) {
  GemmProblem g{};
  g.a_hi = a_hi;
  g.a_lo = a_lo;
  g.a_rows_pad = pad_to(M, 128);
  g.a_ld = a_ld;
  g.a_k0 = a_k0;
  g.b_hi = w.hi;
  g.b_lo = w.lo;
  g.b_rows_pad = w.rows_pad;
  g.b_ld = w.k_pad;
  g.M = M;
  g.N = N;
  g.bn = bn;
  g.tiles_n = (N + bn - 1) / bn;
  g.num_kb = (K + 63) / 64;
  const int tail = K - (g.num_kb - 1) * 64;
  g.last_k4 = (tail + 15) / 16;
  g.epi = epi;
  return g;
}
}  // namespace

extern "C" int sss_encoder_create(sss_encoder_t** out, int device, const sss_encoder_shape_t* shape) {
  SSS_REQUIRE(out && shape, "sss_encoder_create: NULL argument");
  SSS_REQUIRE(shape->in_dim >= 1 && shape->n_layers >= 1 &&
                  shape->out_dim > shape->max_seq_len && shape->max_seq_len >= 1,
              "sss_encoder_create: bad shape");
  SSS_REQUIRE(shape->in_dim <= shape->hidden,
              "The number of input channels is not allowed to be larger than the number of output channels");
  SSS_REQUIRE(shape->in_dim % 8 == 0 && shape->hidden % 8 == 0,
              "sss_encoder_create: in_dim and hidden must be multiples of 8 (16-byte aligned bf16 operand slices)");
  int ndev = 0;
  SSS_CUDA_OK(cudaGetDeviceCount(&ndev));
  SSS_REQUIRE(device >= 0 && device < ndev, "sss_encoder_create: no such CUDA device");
  cudaDeviceProp prop;
  SSS_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  SSS_REQUIRE(prop.major == 10, "libsss_b200 is built for sm_100a only; device is sm_" +
                                    std::to_string(prop.major * 10 + prop.minor));
  sss_encoder* e = new sss_encoder();
  e->device = device;
  e->sh = *shape;
  *out = e;
  return 0;
}

extern "C" int sss_encoder_set_math(sss_encoder_t* e, int math) {
  SSS_REQUIRE(e != nullptr, "sss_encoder_set_math: NULL encoder");
  SSS_REQUIRE(math == SSS_ENCODER_MATH_BF16X3,
              "the encoder's linears run on this library's split-bf16 tcgen05 GEMM (SSS_ENCODER_MATH_BF16X3); the "
              "cuBLAS arithmetics of earlier versions are gone");
  return 0;
}

extern "C" int sss_encoder_get_math(const sss_encoder_t* e) { return e ? SSS_ENCODER_MATH_BF16X3 : -1; }
extern "C" int64_t sss_encoder_stat(const sss_encoder_t* e, int what) {
  if (!e) return -1;
  return what == 0 ? e->launches : what == 1 ? e->host_ns : -1;
}

extern "C" int sss_encoder_destroy(sss_encoder_t* e) {
  if (!e) return 0;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& kv : e->params) cudaFree(kv.second);
  drop_weights(e);
  ws_release(e);
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
  cudaDeviceSynchronize();  // a forward may still be reading the split weights
  drop_weights(e);          // split-bf16 operands are rebuilt on the next forward
  e->gemm_flag = nullptr;
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
    ~Restore() { cudaSetDevice(d); }
  } restore{prev};
  cudaStream_t st = (cudaStream_t)stream;
  if (prepare_weights(e, st)) return 1;
  const auto t_host0 = std::chrono::steady_clock::now();
  struct HostClock {
    sss_encoder* e;
    std::chrono::steady_clock::time_point t0;
    ~HostClock() { e->host_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
  } host_clock{e, t_host0};
  if (ws_begin(e, st)) return 1;
  e->launches = 0;

  // ---- workspace
  const int PP = pad_to(H, 128), GH = pad_to(3 * H, 128), T = PP / 128;
  const int ZDp = pad_to(ZD, 64), Hp = pad_to(H, 64), OUTp = pad_to(OUT, 64);
  const int NQp = pad_to(NQ, 128), NPp = pad_to(NP, 128);
  const int NT = NE + NQ, NTp = pad_to(NT, 128), Bp = pad_to(B, 128);
  const int LDQ = 2 * PP, LDP = 3 * PP + GH;
  const int tiles_out = (OUT + 127) / 128;   // N tiles across the pooling width: partial attention logits per row
  float *Zq = io->z_query, *Zp = io->z_product;  // node embeddings [N, in + layers * hidden]: the caller's or ours
  float *Sq, *Sp, *asq, *adq, *adp, *asp, *Gp, *U, *Bc, *att_part;
  __nv_bfloat16 *zq_hi, *zq_lo, *zp_hi, *zp_lo, *agg_hi, *agg_lo, *u_hi, *u_lo, *c_hi, *c_lo;
  int *prefix, *node_graph, *ranges, *cnt_i, *deg[3], *rowptr[3], *col[3];
  const int n_loop = NQ < NP ? NQ : NP;
  const int64_t e_tot[3] = {bt->e_qp + n_loop, bt->e_pq + n_loop, bt->e_pp};
  const int n_dst[3] = {NP, NQ, NP};
  if ((!Zq && ws_alloc(e, &Zq, (size_t)NQ * ZD)) || (!Zp && ws_alloc(e, &Zp, (size_t)NP * ZD)) ||
      ws_alloc(e, &zq_hi, (size_t)NQp * ZDp) || ws_alloc(e, &zq_lo, (size_t)NQp * ZDp) ||
      ws_alloc(e, &zp_hi, (size_t)NPp * ZDp) || ws_alloc(e, &zp_lo, (size_t)NPp * ZDp) ||
      ws_alloc(e, &u_hi, (size_t)NTp * OUTp) || ws_alloc(e, &u_lo, (size_t)NTp * OUTp) ||
      ws_alloc(e, &c_hi, (size_t)Bp * OUTp) || ws_alloc(e, &c_lo, (size_t)Bp * OUTp) ||
      ws_alloc(e, &Sq, (size_t)NQ * LDQ) || ws_alloc(e, &Sp, (size_t)NP * LDP) || ws_alloc(e, &asq, (size_t)NQ * T * 2) ||
      ws_alloc(e, &adq, (size_t)NQ * T * 2) || ws_alloc(e, &adp, (size_t)NP * T * 2) || ws_alloc(e, &asp, (size_t)NP * T * 2) ||
      ws_alloc(e, &Gp, (size_t)NP * H) || ws_alloc(e, &agg_hi, (size_t)NPp * Hp) || ws_alloc(e, &agg_lo, (size_t)NPp * Hp) ||
      ws_alloc(e, &U, (size_t)NT * OUT) || ws_alloc(e, &Bc, (size_t)B * OUT) ||
      ws_alloc(e, &att_part, (size_t)NT * tiles_out * 2) || ws_alloc(e, &prefix, NP + 1) || ws_alloc(e, &node_graph, NT + 1) ||
      ws_alloc(e, &ranges, (size_t)B * 4) || ws_alloc(e, &cnt_i, NP + 1))
    return 1;
  for (int j = 0; j < 3; ++j)
    if (ws_alloc(e, &deg[j], n_dst[j] + 1) || ws_alloc(e, &rowptr[j], n_dst[j] + 1) || ws_alloc(e, &col[j], e_tot[j] + 1))
      return 1;
  if (nonfinite) SSS_CUDA_OK(cudaMemsetAsync(nonfinite, 0, sizeof(int32_t), st));
  if ((IN % 16) || (H % 16) || (OUT % 16)) {
    // K slices that are not whole 16-wide MMA steps read a few operand columns beyond their end (against zero weight
    // columns): those must be finite, so shapes like that (never the reference's 768 / 800 / 1600) start from zeros
    SSS_CUDA_OK(cudaMemsetAsync(zq_hi, 0, (size_t)NQp * ZDp * 2, st));
    SSS_CUDA_OK(cudaMemsetAsync(zq_lo, 0, (size_t)NQp * ZDp * 2, st));
    SSS_CUDA_OK(cudaMemsetAsync(zp_hi, 0, (size_t)NPp * ZDp * 2, st));
    SSS_CUDA_OK(cudaMemsetAsync(zp_lo, 0, (size_t)NPp * ZDp * 2, st));
    SSS_CUDA_OK(cudaMemsetAsync(u_hi, 0, (size_t)NTp * OUTp * 2, st));
    SSS_CUDA_OK(cudaMemsetAsync(u_lo, 0, (size_t)NTp * OUTp * 2, st));
    SSS_CUDA_OK(cudaMemsetAsync(c_hi, 0, (size_t)Bp * OUTp * 2, st));
    SSS_CUDA_OK(cudaMemsetAsync(c_lo, 0, (size_t)Bp * OUTp * 2, st));
  }

  // ---- ingest + graph structure
  if (run_gnn) {
    SSS_REQUIRE(bt->x_query && bt->x_product, "sss_encoder_forward: the GNN stage needs x_query / x_product");
    ingest_kernel<<<NQ + NP, 256, 0, st>>>(bt->x_query, NQ, bt->x_product, NP, IN, IN, Zq, Zp, ZD, zq_hi, zq_lo, zp_hi,
                                           zp_lo, ZDp, nonfinite);
  } else {  // pooling alone: the caller's node embeddings become the GEMM operands
    ingest_kernel<<<NQ + NP, 256, 0, st>>>(Zq, NQ, Zp, NP, ZD, ZD, Zq, Zp, ZD, zq_hi, zq_lo, zp_hi, zp_lo, ZDp, nullptr);
  }
  PrepArgs pa{};
  pa.run_csr = run_gnn ? 1 : 0;
  pa.csr[0] = CsrJob{bt->qp_src, bt->qp_dst, bt->e_qp, NP, 1, n_loop, deg[0], rowptr[0], col[0]};  // dst = product
  pa.csr[1] = CsrJob{bt->pq_src, bt->pq_dst, bt->e_pq, NQ, 1, n_loop, deg[1], rowptr[1], col[1]};  // dst = query
  pa.csr[2] = CsrJob{bt->pp_src, bt->pp_dst, bt->e_pp, NP, 0, 0, deg[2], rowptr[2], col[2]};       // dst = product
  pa.cnt = bt->product_cnt;
  pa.product_batch = bt->product_batch;
  pa.query_batch = bt->query_batch;
  pa.n_p = NP;
  pa.n_q = NQ;
  pa.n_e = NE;
  pa.n_graphs = B;
  pa.prefix = prefix;
  pa.node_graph = node_graph;
  pa.ranges = ranges;
  pa.cnt_i = cnt_i;
  graph_prep_kernel<<<4, 1024, 0, st>>>(pa);
  e->launches += 2;

  // ---- HeteroGGNN layers (model/gnn.py:64-81)
  for (int l = 0; run_gnn && l < L; ++l) {
    const int cin = l == 0 ? IN : H;
    const int off = l == 0 ? 0 : IN + (l - 1) * H;
    const int off_next = IN + l * H;
    const std::string pre = "gnn.convs." + std::to_string(l) + ".convs.";
    const std::string eqp = pre + "query__clicks__product.", epq = pre + "product__clicked by__query.",
                      epp = pre + "product__to__product.";
    const float *a_qp_src = P(e, eqp + "att_src", H), *a_qp_dst = P(e, eqp + "att_dst", H), *b_qp = P(e, eqp + "bias", H),
                *a_pq_src = P(e, epq + "att_src", H), *a_pq_dst = P(e, epq + "att_dst", H), *b_pq = P(e, epq + "bias", H),
                *b_ih = P(e, epp + "rnn.bias_ih", 3 * H), *b_hh = P(e, epp + "rnn.bias_hh", 3 * H);
    if (!a_qp_src || !a_qp_dst || !b_qp || !a_pq_src || !a_pq_dst || !b_pq || !b_ih || !b_hh) return 1;
    // one launch: query side [S_qp | T_pq] and product side [T_qp | S_pq | M | gh], attention logits as partials
    GemmProblem g2[2];
    g2[0] = make_problem(zq_hi, zq_lo, NQ, ZDp, off, cin, e->w_q[l], LDQ, 128, EPI_ATT);
    g2[0].C = Sq;
    g2[0].ldc = LDQ;
    g2[0].att[0] = a_qp_src; g2[0].att_out[0] = asq;   // a_s of q -> p
    g2[0].att[1] = a_pq_dst; g2[0].att_out[1] = adq;   // a_d of p -> q
    g2[0].att_tiles_per_part = T;
    g2[0].att_width = H;
    g2[1] = make_problem(zp_hi, zp_lo, NP, ZDp, off, cin, e->w_p[l], LDP, 128, EPI_ATT);
    g2[1].C = Sp;
    g2[1].ldc = LDP;
    g2[1].att[0] = a_qp_dst; g2[1].att_out[0] = adp;   // a_d of q -> p
    g2[1].att[1] = a_pq_src; g2[1].att_out[1] = asp;   // a_s of p -> q
    g2[1].att_tiles_per_part = T;
    g2[1].att_width = H;
    if (launch_gemm_bf16x3(g2, 2, e->gemm_flag, st)) return 1;
    MpArgs ma{};
    ma.n_p = NP; ma.n_q = NQ; ma.H = H; ma.parts = 2 * T;   // (row, 128-column sub-tile, epilogue warpgroup)
    ma.qp_rowptr = rowptr[0]; ma.qp_col = col[0]; ma.pq_rowptr = rowptr[1]; ma.pq_col = col[1];
    ma.pp_rowptr = rowptr[2]; ma.pp_col = col[2];
    ma.Sq = Sq; ma.ldsq = LDQ; ma.Sp = Sp; ma.ldsp = LDP; ma.PP = PP;
    ma.asq = asq; ma.adq = adq; ma.adp = adp; ma.asp = asp;
    ma.b_qp = b_qp; ma.b_pq = b_pq;
    ma.Gp = Gp;
    ma.agg_hi = agg_hi; ma.agg_lo = agg_lo; ma.ld_agg = Hp;
    ma.zq_next = Zq + off_next; ma.ldz = ZD;
    ma.zq_hi = zq_hi; ma.zq_lo = zq_lo; ma.ldh = ZDp; ma.k0_next = off_next;
    message_passing_kernel<<<NP + NQ, 256, 0, st>>>(ma);
    // GatedGraphConv's GRUCell on the aggregated messages, HeteroConv sum with the GAT branch, relu: one GEMM
    GemmProblem g3 = make_problem(agg_hi, agg_lo, NP, Hp, 0, H, e->w_ih[l], ((H + 31) / 32) * 96, 96, EPI_GRU);
    g3.C = Zp + off_next;
    g3.ldc = ZD;
    g3.gru_gh = Sp + 3 * PP;
    g3.gru_gh_ld = LDP;
    g3.gru_b_ih = b_ih;
    g3.gru_b_hh = b_hh;
    g3.gru_x = Zp + off;
    g3.gru_x_ld = ZD;
    g3.gru_in_w = cin;
    g3.gru_gp = Gp;
    g3.gru_H = H;
    g3.out_hi = zp_hi;
    g3.out_lo = zp_lo;
    g3.out_ld = ZDp;
    g3.out_k0 = off_next;
    if (launch_gemm_bf16x3(&g3, 1, e->gemm_flag, st)) return 1;
    e->launches += 3;
  }
  if (!run_pool) {
    SSS_CUDA_OK(cudaGetLastError());
    return ws_end(e, st);
  }

  // ---- PositionalAttentionPooling (model/gnn.py:193-217)
  const float *bq = P(e, "pooling.query_lin.bias", LIN), *bp = P(e, "pooling.product_lin.bias", LIN),
              *pe = P(e, "pooling.positional_emb.weight", (int64_t)MSL * MSL),
              *bn = P(e, "pooling.node_emb_lin.bias", OUT), *wa = P(e, "pooling.att_lin.weight", OUT);
  if (!bq || !bp || !pe || !bn || !wa) return 1;
  GemmProblem gp2[2];
  for (int s = 0; s < 2; ++s) {  // 0: products (rows [0, NE) of U after repeat_interleave), 1: queries (rows NE ..)
    const bool prod = s == 0;
    gp2[s] = make_problem(prod ? zp_hi : zq_hi, prod ? zp_lo : zq_lo, prod ? NP : NQ, ZDp, 0, ZD, prod ? e->w_poolp : e->w_poolq,
                          OUT, 128, EPI_POOL);
    gp2[s].bias = prod ? bp : bq;
    gp2[s].pool_is_product = prod ? 1 : 0;
    gp2[s].pool_row0 = NE;
    gp2[s].pool_lin_w = LIN;
    gp2[s].pool_msl = MSL;
    gp2[s].pool_prefix = prefix;
    gp2[s].pool_pos = prod ? bt->product_pos : bt->query_pos;
    gp2[s].pool_pe = pe;
    gp2[s].pool_U = U;
    gp2[s].out_hi = u_hi;
    gp2[s].out_lo = u_lo;
    gp2[s].out_ld = OUTp;
  }
  if (launch_gemm_bf16x3(gp2, 2, e->gemm_flag, st)) return 1;
  const int col_blocks = std::max(1, std::min(8, (OUT + 1023) / 1024));  // 1024 columns per block: one pass
  graph_mean_hilo_kernel<<<dim3(B, col_blocks), 256, 0, st>>>(U, OUT, ranges, c_hi, c_lo, OUTp);
  GemmProblem gc = make_problem(c_hi, c_lo, B, OUTp, 0, OUT, e->w_coarse, OUT, 128, EPI_STORE);
  gc.C = Bc;
  gc.ldc = OUT;
  if (launch_gemm_bf16x3(&gc, 1, e->gemm_flag, st)) return 1;
  GemmProblem gn = make_problem(u_hi, u_lo, NT, OUTp, 0, OUT, e->w_node, OUT, 128, EPI_ATTPOOL);
  gn.bias = bn;
  gn.ap_bc = Bc;
  gn.ap_node_graph = node_graph;
  gn.ap_w = wa;
  gn.ap_out = att_part;
  if (launch_gemm_bf16x3(&gn, 1, e->gemm_flag, st)) return 1;
  graph_weighted_mean_kernel<<<dim3(B, col_blocks), 256, 1024 * sizeof(float), st>>>(U, OUT, ranges, att_part, 2 * tiles_out, out);
  e->launches += 5;
  SSS_CUDA_OK(cudaGetLastError());
  return ws_end(e, st);
}

extern "C" int sss_binarize_head(const float* x, const float* W, const float* b, int64_t n, int in_dim, int out_dim,
                                 float* out, int device, void* stream) {
  SSS_REQUIRE(x && W && b && out, "sss_binarize_head: NULL buffer");
  SSS_REQUIRE(n >= 0 && in_dim >= 1 && out_dim >= 1, "sss_binarize_head: bad shape");
  if (n == 0) return 0;
  int prev = 0;
  SSS_CUDA_OK(cudaGetDevice(&prev));
  SSS_CUDA_OK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  // out = sign(x W^T + b) on the same split-bf16 tcgen05 GEMM as the encoder's linears (sign in the epilogue)
  const int k_pad = pad_to(in_dim, 64), m_pad = pad_to((int)n, 128), n_pad = pad_to(out_dim, 128);
  const size_t a_elems = (size_t)m_pad * k_pad, b_elems = (size_t)n_pad * k_pad;
  uint16_t* buf = nullptr;
  int* flag = nullptr;
  cudaError_t err = cudaMallocAsync((void**)&buf, (2 * a_elems + 2 * b_elems) * 2 + 256, st);
  int rc = err == cudaSuccess ? 0 : 1;
  if (!rc) {
    flag = (int*)(buf + 2 * a_elems + 2 * b_elems);
    if (cudaMemsetAsync(flag, 0, sizeof(int), st) != cudaSuccess) rc = 1;
  }
  if (!rc) rc = launch_split_bf16(x, (int)n, in_dim, in_dim, 0, buf, buf + a_elems, m_pad, k_pad, st);
  if (!rc) rc = launch_split_bf16(W, out_dim, in_dim, in_dim, 0, buf + 2 * a_elems, buf + 2 * a_elems + b_elems, n_pad, k_pad, st);
  if (!rc) {
    SplitW w;
    w.hi = buf + 2 * a_elems;
    w.lo = buf + 2 * a_elems + b_elems;
    w.rows_pad = n_pad;
    w.k_pad = k_pad;
    GemmProblem g = make_problem(buf, buf + a_elems, (int)n, k_pad, 0, in_dim, w, out_dim, 128, EPI_STORE);
    g.C = out;
    g.ldc = out_dim;
    g.bias = b;
    g.sign_out = 1;
    rc = launch_gemm_bf16x3(&g, 1, flag, st);
  }
  if (buf) cudaFreeAsync(buf, st);
  cudaSetDevice(prev);
  if (err != cudaSuccess) set_error(std::string("sss_binarize_head: ") + cudaGetErrorString(err));
  return rc;
}
