// api.cu — the C ABI of libsss_b200.so (include/sss_b200.h): handles, workspaces and the wave driver
// that strings scan -> expand -> refine -> emit together on the caller's stream.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/sss_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace sss {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <typename T>
static int dev_alloc(T** p, size_t count) {
  void* v = nullptr;
  SSS_CUDA_OK(cudaMalloc(&v, std::max<size_t>(count, 1) * sizeof(T)));
  *p = (T*)v;
  return 0;
}
template <typename T>
static void dev_free(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

// Workspace of one search call; grows monotonically.
struct Workspace {
  int64_t nq_pad = 0;
  int cap = 0, d = 0, d_pad = 0, k = 0;
  float *thr = nullptr, *margin = nullptr, *qn2 = nullptr, *q_f32 = nullptr, *q_keep = nullptr, *out_D = nullptr;
  uint32_t *cnt = nullptr, *nret = nullptr, *skip_list = nullptr, *skip_cnt = nullptr;
  uint64_t* cand = nullptr;
  int* flags = nullptr;       // [0] overflow, [1] kernel watchdog
  int* host_flags = nullptr;  // pinned copy of flags (the one status word a search reads back)
  void* q_bf16 = nullptr;
  int64_t* out_I = nullptr;
  HitRecord* rec = nullptr;
  uint32_t* rec_cnt = nullptr;
  float* cmax = nullptr;      // bootstrap chunk maxima [n_chunks, nq_pad]
  size_t cmax_elems = 0;
  size_t rec_entries = 0;
  int n_regions = 0;
  int64_t out_elems = 0;
  uint64_t generation = 0;    // bumped whenever a buffer is reallocated (captured graphs hold raw pointers)

  int ensure(int64_t nq_pad_, int cap_, int d_, int d_pad_, int64_t out_elems_) {
    if (nq_pad_ > nq_pad || cap_ > cap || d_ != d || d_pad_ != d_pad) {
      release_query_side();
      ++generation;
      nq_pad = std::max(nq_pad_, nq_pad);
      cap = std::max(cap_, cap);
      d = d_;
      d_pad = d_pad_;
      if (dev_alloc(&thr, nq_pad) || dev_alloc(&margin, nq_pad) || dev_alloc(&qn2, nq_pad) || dev_alloc(&cnt, nq_pad) ||
          dev_alloc(&nret, nq_pad) || dev_alloc(&skip_list, nq_pad) || dev_alloc(&skip_cnt, 2) ||
          dev_alloc(&cand, (size_t)nq_pad * cap) || dev_alloc(&q_f32, (size_t)nq_pad * d) ||
          dev_alloc(&q_keep, (size_t)nq_pad * d))
        return 1;
      void* v = nullptr;
      SSS_CUDA_OK(cudaMalloc(&v, (size_t)nq_pad * d_pad * 2));
      q_bf16 = v;
    }
    if (!flags) {
      if (dev_alloc(&flags, 2)) return 1;
      SSS_CUDA_OK(cudaHostAlloc((void**)&host_flags, 2 * sizeof(int), cudaHostAllocDefault));
      ++generation;
    }
    if (out_elems_ > out_elems) {
      dev_free(out_D);
      dev_free(out_I);
      ++generation;
      out_elems = out_elems_;
      if (dev_alloc(&out_D, out_elems) || dev_alloc(&out_I, out_elems)) return 1;
    }
    return 0;
  }
  int ensure_records(int n_regions_, int rec_cap) {
    size_t need = (size_t)n_regions_ * rec_cap;
    if (need > rec_entries) {
      dev_free(rec);
      ++generation;
      rec_entries = need;
      if (dev_alloc(&rec, rec_entries)) return 1;
    }
    if (n_regions_ > n_regions) {
      dev_free(rec_cnt);
      ++generation;
      n_regions = n_regions_;
      if (dev_alloc(&rec_cnt, n_regions)) return 1;
    }
    return 0;
  }
  int ensure_cmax(size_t need) {
    if (need > cmax_elems) {
      dev_free(cmax);
      ++generation;
      cmax_elems = need;
      if (dev_alloc(&cmax, need)) return 1;
    }
    return 0;
  }
  void release_query_side() {
    dev_free(thr); dev_free(margin); dev_free(qn2); dev_free(cnt); dev_free(nret); dev_free(skip_list); dev_free(skip_cnt);
    dev_free(cand); dev_free(q_f32); dev_free(q_keep);
    if (q_bf16) cudaFree(q_bf16);
    q_bf16 = nullptr;
  }
  void release() {
    release_query_side();
    dev_free(flags); dev_free(out_D); dev_free(out_I); dev_free(rec); dev_free(rec_cnt); dev_free(cmax);
    if (host_flags) cudaFreeHost(host_flags);
    host_flags = nullptr;
    cmax_elems = 0;
    nq_pad = 0; cap = 0; rec_entries = 0; n_regions = 0; out_elems = 0;
    ++generation;
  }
  SelectState state() const {
    SelectState s;
    s.thr = thr; s.cnt = cnt; s.nret = nret; s.cand = cand; s.margin = margin; s.qn2 = qn2; s.skip_list = skip_list; s.skip_cnt = skip_cnt;
    s.overflow = flags;
    s.cap = cap;
    return s;
  }
};

// A growable row store: fixed-order-normalised fp32 rows (rescoring / fp32 scan) and, when the tensor
// path applies (inner product, d <= 4096), a zero-padded bf16 copy laid out for TMA (row pitch d_pad*2 bytes).
struct RowStore {
  float* f32 = nullptr;
  void* bf16 = nullptr;
  unsigned int* maxnorm2 = nullptr;  // device [4]: bits of max ||row||^2, max ||row - bf16(row)||^2, max ||row||_4^4
  int64_t n = 0, cap_rows = 0;
  int ensure(int64_t rows, int d, int d_pad, bool want_bf16, cudaStream_t st) {
    if (!maxnorm2) {
      if (dev_alloc(&maxnorm2, 4)) return 1;
      SSS_CUDA_OK(cudaMemsetAsync(maxnorm2, 0, 4 * sizeof(unsigned int), st));
    }
    if (rows <= cap_rows) return 0;
    int64_t new_cap = std::max<int64_t>(rows, cap_rows + cap_rows / 2);
    new_cap = (new_cap + 127) / 128 * 128;
    float* nf = nullptr;
    if (dev_alloc(&nf, (size_t)new_cap * d)) return 1;
    if (n > 0) SSS_CUDA_OK(cudaMemcpyAsync(nf, f32, (size_t)n * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    void* nb = nullptr;
    if (want_bf16) {
      SSS_CUDA_OK(cudaMalloc(&nb, (size_t)new_cap * d_pad * 2));
      if (n > 0) SSS_CUDA_OK(cudaMemcpyAsync(nb, bf16, (size_t)n * d_pad * 2, cudaMemcpyDeviceToDevice, st));
    }
    SSS_CUDA_OK(cudaStreamSynchronize(st));
    dev_free(f32);
    if (bf16) cudaFree(bf16);
    f32 = nf;
    bf16 = nb;
    cap_rows = new_cap;
    return 0;
  }
  void release() {
    dev_free(f32);
    if (bf16) cudaFree(bf16);
    bf16 = nullptr;
    dev_free(maxnorm2);
    n = cap_rows = 0;
  }
};

}  // namespace sss

using namespace sss;

// Environment knobs, read ONCE when a handle is created (tuning and A/B runs; production needs none of them).
struct Tuning {
  int growth10 = 0;      // SSS_WAVE_GROWTH: wave growth factor x10 after the bootstrap pass (0 = default)
  int64_t first = 0;     // SSS_WAVE_FIRST: rows of the first wave (0 = default)
  bool no_bootstrap = false, no_lazy = false, no_graph = false;
  int variant = 0;       // SSS_SCAN_VARIANT: ss | ts | 2cta (0 = automatic)
  static Tuning from_env() {
    Tuning t;
    if (const char* g = getenv("SSS_WAVE_GROWTH")) t.growth10 = (int)std::max<int64_t>(11, atoll(g));
    if (const char* f = getenv("SSS_WAVE_FIRST")) t.first = std::max<int64_t>(512, atoll(f) / 512 * 512);
    auto on = [](const char* n) { const char* v = getenv(n); return v && v[0] == '1'; };
    t.no_bootstrap = on("SSS_NO_BOOTSTRAP");
    t.no_lazy = on("SSS_NO_LAZY");
    t.no_graph = on("SSS_NO_GRAPH");
    if (const char* v = getenv("SSS_SCAN_VARIANT")) t.variant = v[0] == 's' ? 1 : v[0] == 't' ? 2 : v[0] == '2' ? 3 : 0;
    return t;
  }
};

// One captured search: every launch from query staging to emit as a CUDA graph.  A step of the headline workload is
// ~20 dependent kernels of 3-800 us; replayed as a graph they cost one host call and the launch gaps shrink to the
// hardware's node-to-node latency.  Only two pointers differ between calls — the caller's query buffer (read by
// prep_queries_kernel alone) and the output buffers (written by emit_kernel alone) — and those two kernel nodes are
// re-bound with cudaGraphExecKernelNodeSetParams before every replay.
struct SearchGraph {
  int64_t nq = 0;
  int k = 0, mode = 0;
  uint64_t epoch = 0, ws_gen = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t prep = nullptr, emit = nullptr;
  int64_t kernels = 0, waves = 0, variant = 0;
  uint64_t last_use = 0;
  void destroy() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    exec = nullptr;
    graph = nullptr;
  }
};

struct sss_index {
  int device = 0, d = 0, d_pad = 0, metric = 0, num_sms = 148;
  int64_t id_offset = 0;
  bool tensor_ok = false;
  int rec_boost = 1;  // 4 once a search overflowed a record sub-region (sticky: such data would do it again)
  RowStore rows;      // the added rows
  RowStore sums;      // per-segment sums (reduce == SUM)
  int reduce = 0;
  int64_t n_seg = 0;
  int64_t max_seg_len = 1;
  int64_t* seg_off = nullptr;  // device
  int32_t* row_seg = nullptr;  // device
  Workspace ws;
  Tuning tune;
  uint64_t epoch = 1;          // bumped by add / set_segments / rec_boost changes: invalidates captured graphs
  std::vector<SearchGraph> graphs;
  uint64_t graph_clock = 0;
  cudaStream_t cap_stream = nullptr;  // capture happens here (the caller's stream may be the legacy default stream)
  int64_t stat_kernels = 0, stat_waves = 0, stat_reruns = 0, stat_overflow_reason = 0, stat_variant = 0, stat_graph = 0;
  // optional scan-kernel timing (CUDA events on the launching stream around every scan launch; eager launches)
  bool profile = false;
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  double scan_us = 0.0;
  int64_t scan_launches = 0;
  unsigned long long* dbg = nullptr;  // device [12], see RefineArgs::debug
  unsigned long long dbg_host[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // Binary flavour (sss_binary_index wraps one of these): rows.f32 holds the PACKED codes (pitch = 4 * d bytes per
  // row), rows.bf16 their +-1 E4M3 expansion (2 * d_pad bytes per row); scores are integers, nothing is re-scored.
  bool binary = false;
  int b_nbits = 0, b_nbytes = 0;
  int64_t q_row_bytes() const { return binary ? b_nbytes : (int64_t)d * 4; }
  void drop_graphs() {
    for (auto& g : graphs) g.destroy();
    graphs.clear();
  }
};

extern "C" const char* sss_last_error(void) { return g_err.c_str(); }
extern "C" int sss_version(void) { return 200; }
extern "C" int sss_built_for_sm(void) { return 100; }

extern "C" int sss_index_create(sss_index_t** out, int device, int d, int metric, int64_t id_offset) {
  SSS_REQUIRE(out != nullptr, "sss_index_create: out is NULL");
  SSS_REQUIRE(d >= 1, "sss_index_create: d must be >= 1");
  SSS_REQUIRE(metric == SSS_METRIC_IP || metric == SSS_METRIC_L2, "Unregnozed metric");
  int ndev = 0;
  SSS_CUDA_OK(cudaGetDeviceCount(&ndev));
  SSS_REQUIRE(device >= 0 && device < ndev, "sss_index_create: no such CUDA device");
  cudaDeviceProp prop;
  SSS_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  SSS_REQUIRE(prop.major == 10, "libsss_b200 is built for sm_100a only; device is sm_" +
                                    std::to_string(prop.major * 10 + prop.minor));
  sss_index* ix = new sss_index();
  ix->device = device;
  ix->d = d;
  // (L2 on the tensor path keeps -||x||^2 / 2 in two extra bf16 columns, hi + lo: prep.cu add_rows_kernel)
  ix->d_pad = (d + (metric == SSS_METRIC_L2 ? 2 : 0) + 63) / 64 * 64;
  ix->metric = metric;
  ix->id_offset = id_offset;
  ix->num_sms = prop.multiProcessorCount;
  ix->tensor_ok = ix->d_pad <= 4096;
  ix->tune = Tuning::from_env();
  *out = ix;
  return 0;
}

extern "C" int sss_index_destroy(sss_index_t* ix) {
  if (!ix) return 0;
  DeviceGuard g(ix->device);
  if (g.ok) {
    cudaDeviceSynchronize();  // nothing of this handle may still be in flight when its buffers go
    ix->drop_graphs();
    ix->rows.release();
    ix->sums.release();
    dev_free(ix->seg_off);
    dev_free(ix->row_seg);
    dev_free(ix->dbg);
    ix->ws.release();
    for (cudaEvent_t e : ix->ev) cudaEventDestroy(e);
    if (ix->cap_stream) cudaStreamDestroy(ix->cap_stream);
  }
  delete ix;
  return g.ok ? 0 : 1;
}

extern "C" int64_t sss_index_ntotal(const sss_index_t* ix) { return ix ? ix->rows.n : 0; }
extern "C" int sss_index_dim(const sss_index_t* ix) { return ix ? ix->d : 0; }
extern "C" int64_t sss_index_stat(const sss_index_t* ix, int what) {
  if (!ix) return -1;
  switch (what) {
    case 0: return ix->stat_kernels;
    case 1: return ix->stat_waves;
    case 2: return ix->stat_reruns;
    case 3: return (int64_t)(ix->scan_us * 1000.0);  // scan-kernel time of the last search, ns (profiling on)
    case 4: return ix->scan_launches;
    case 5: return (int64_t)ix->dbg_host[0];  // candidates entering refine (profiling on)
    case 6: return (int64_t)ix->dbg_host[1];  // rows re-scored
    case 7: return (int64_t)ix->dbg_host[2];  // sessions sorted
    case 8: return (int64_t)ix->dbg_host[3];  // refine invocations (queries x waves with new candidates)
#ifdef SSS_EXPERIMENT
    case 9: case 10: case 11: case 12: case 13: case 14: case 15:
      return (int64_t)ix->dbg_host[4 + (what - 9)];  // cycles of refine phase (what - 9), summed over invocations
#endif
    case 24: return ix->stat_overflow_reason;  // bit mask of what overflowed in the last rerun (select.cu)
    case 25: return ix->stat_variant;  // scan of the last search: 0 fp32, 1 SS, 2 TS, 3 pair (2-CTA), 4 K-loop pair
    case 26: return ix->stat_graph;    // 1 when the last search replayed a captured CUDA graph
    default: return -1;
  }
}
extern "C" int sss_index_set_profiling(sss_index_t* ix, int on) {
  SSS_REQUIRE(ix != nullptr, "sss_index_set_profiling: NULL index");
  ix->profile = on != 0;
  return 0;
}

extern "C" int sss_index_add(sss_index_t* ix, const float* rows, int64_t n, int rows_on_device, int norm_mode,
                             void* stream) {
  SSS_REQUIRE(ix != nullptr, "sss_index_add: NULL index");
  SSS_REQUIRE(n >= 0, "sss_index_add: negative row count");
  SSS_REQUIRE(norm_mode >= 0 && norm_mode <= 3, "sss_index_add: unknown norm_mode");
  if (n == 0) return 0;
  SSS_REQUIRE(rows != nullptr, "sss_index_add: NULL rows");
  SSS_REQUIRE(ix->rows.n + n < 0xFFFFFFF0ll, "sss_index_add: more than 2^32 rows in one shard");
  DeviceGuard g(ix->device);
  SSS_REQUIRE(g.ok, "sss_index_add: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  ix->epoch += 1;
  ix->drop_graphs();
  if (ix->rows.ensure(ix->rows.n + n, ix->d, ix->d_pad, ix->tensor_ok, st)) return 1;
  const float* src = rows;
  float* staged = nullptr;
  if (!rows_on_device) {
    if (dev_alloc(&staged, (size_t)n * ix->d)) return 1;
    if (cudaMemcpyAsync(staged, rows, (size_t)n * ix->d * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) {
      cudaFree(staged);
      set_error("sss_index_add: host to device copy failed");
      return 1;
    }
    src = staged;
  }
  int rc = launch_add_rows(src, n, ix->d, ix->d_pad, norm_mode, ix->rows.f32, ix->rows.bf16, ix->rows.n,
                           ix->rows.maxnorm2, st, (ix->tensor_ok && ix->metric == SSS_METRIC_L2) ? 1 : 0);
  if (staged) {
    cudaStreamSynchronize(st);
    cudaFree(staged);
  }
  if (rc) return rc;
  ix->rows.n += n;
  ix->reduce = 0;  // segments must be re-declared after adding rows
  ix->n_seg = 0;
  return 0;
}

// All of its work is enqueued on `stream` — the stream the rows were added on, or one ordered after it — and the call
// returns after that stream has drained (the segment offsets are host memory and the summed rows are built here).
extern "C" int sss_index_set_segments(sss_index_t* ix, const int64_t* seg_off, int64_t n_seg, int reduce, void* stream) {
  SSS_REQUIRE(ix != nullptr, "sss_index_set_segments: NULL index");
  SSS_REQUIRE(reduce >= 0 && reduce <= 2, "sss_index_set_segments: unknown reduce");
  DeviceGuard g(ix->device);
  SSS_REQUIRE(g.ok, "sss_index_set_segments: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  if (reduce == SSS_REDUCE_NONE) {
    ix->epoch += 1;
    ix->drop_graphs();
    ix->reduce = 0;
    ix->n_seg = 0;
    return 0;
  }
  // validate everything before touching the index: a rejected call leaves it exactly as it was
  SSS_REQUIRE(seg_off != nullptr && n_seg >= 1, "sss_index_set_segments: need seg_off[n_seg+1]");
  SSS_REQUIRE(seg_off[0] == 0 && seg_off[n_seg] == ix->rows.n, "sss_index_set_segments: seg_off must span [0, ntotal]");
  SSS_REQUIRE(reduce != SSS_REDUCE_SUM || ix->metric == SSS_METRIC_IP,
              "reduce = sum is defined for the inner-product metric only (a sum of distances is not a distance to a sum)");
  int64_t max_len = 1;
  for (int64_t s = 0; s < n_seg; ++s) {
    SSS_REQUIRE(seg_off[s + 1] >= seg_off[s], "sss_index_set_segments: seg_off must be non-decreasing");
    max_len = std::max(max_len, seg_off[s + 1] - seg_off[s]);
  }
  int64_t* new_off = nullptr;
  int32_t* new_map = nullptr;
  float* tmp = nullptr;
  RowStore new_sums;
  int rc = dev_alloc(&new_off, n_seg + 1) || dev_alloc(&new_map, ix->rows.n);
  if (!rc && cudaMemcpyAsync(new_off, seg_off, (size_t)(n_seg + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st) != cudaSuccess) {
    set_error("sss_index_set_segments: host to device copy failed");
    rc = 1;
  }
  if (!rc) rc = launch_row_seg(new_off, n_seg, new_map, st);
  if (!rc && reduce == SSS_REDUCE_SUM) {
    // linearity: sum_r <q, x_r> = <q, sum_r x_r>; build the summed rows once, then search them as rows
    rc = new_sums.ensure(n_seg, ix->d, ix->d_pad, ix->tensor_ok, st) || dev_alloc(&tmp, (size_t)n_seg * ix->d);
    if (!rc) rc = launch_segment_sum(ix->rows.f32, new_off, n_seg, ix->d, tmp, st);
    if (!rc) rc = launch_add_rows(tmp, n_seg, ix->d, ix->d_pad, 0, new_sums.f32, new_sums.bf16, 0, new_sums.maxnorm2, st);
    if (!rc) new_sums.n = n_seg;
  }
  if (cudaStreamSynchronize(st) != cudaSuccess && !rc) {
    set_error("sss_index_set_segments: CUDA failure while building the segment map");
    rc = 1;
  }
  dev_free(tmp);
  if (rc) {
    dev_free(new_off);
    dev_free(new_map);
    new_sums.release();
    return 1;
  }
  ix->epoch += 1;
  ix->drop_graphs();
  dev_free(ix->seg_off);
  dev_free(ix->row_seg);
  ix->sums.release();
  ix->seg_off = new_off;
  ix->row_seg = new_map;
  ix->sums = new_sums;
  ix->max_seg_len = max_len;
  ix->reduce = reduce;
  ix->n_seg = n_seg;
  return 0;
}

namespace sss {

// Row-ordered scan waves.  The first wave has no threshold, so it must fit the candidate lists; later
// waves grow geometrically (each yields ~k*(growth-1) candidates per query on exchangeable data).
static std::vector<int64_t> make_waves(int64_t n_rows, int cap, int k, bool safe, bool dense_groups,
                                       int64_t bootstrap_rows, const Tuning& tune, bool cautious = false,
                                       bool quad = false) {
  std::vector<int64_t> ends;
  if (bootstrap_rows > 0) {  // thresholds come from a chunk-max pass over [0, bootstrap_rows): re-scan those rows first
    // Between waves the lazy refine costs ~30 us per wave almost independently of the candidate volume
    // (profiles/r02_wave_growth.md: 10M rows x 1000 queries, x2 / 8 waves 2.46 ms, x3 / 5 waves 2.39 ms, x4 2.56 ms:
    // above x3 the record sub-regions of a (query, pair, warpgroup) start to overflow their 16 records).
    // cautious: the second attempt of a batch whose candidate volume outgrew a refine limit (near-duplicate rows:
    // many rows inside the 2 * margin band) — a third of the candidates per wave, before the 2048-row safe schedule
    const int64_t growth10 = cautious ? 14 : tune.growth10 ? tune.growth10 : 30;
    int64_t e = cautious ? bootstrap_rows / 2 : tune.first ? tune.first : bootstrap_rows;
    e = std::min(e, n_rows);
    ends.push_back(e);
    while (e < n_rows) {
      int64_t nx = (e * growth10 / 10 + 511) / 512 * 512;
      // do not leave a short tail wave: fold anything below a quarter wave into this one
      if (n_rows - nx < (nx - e) / 4) nx = n_rows;
      e = std::min(n_rows, nx);
      ends.push_back(e);
    }
    return ends;
  }
  int64_t first = std::max<int64_t>(128, (std::min<int64_t>(cap / 2, cap - k) / 128) * 128);
  if (first > 2048) first = 2048;
  // without session grouping every row is its own group: a shorter threshold-less first wave keeps the
  // common case inside the small refine instantiation (<= 768 groups)
  if (!safe && !dense_groups && k <= 256) first = 512;
  int64_t e = std::min(n_rows, first);
  ends.push_back(e);
  while (e < n_rows) {
    // rows of one session pass the filter together, so session-reduced searches see several candidate rows
    // per new session: keep those waves to a doubling
    int64_t step = safe ? first : ((e < 262144 && !dense_groups) || quad) ? 3 * e : e;
    e = std::min(n_rows, e + step);
    ends.push_back(e);
  }
  return ends;
}

struct BatchArgs {
  const float* q;
  int64_t nq;
  int k, mode;
  bool q_on_device, out_on_device;
  float* D;
  int64_t* I;
  int* status_out = nullptr;  // asynchronous form: no status read-back, no re-run; emit ORs the status word in here
};

// Everything a search pass needs that does not depend on the attempt: operands, plan, tensor maps.
struct SearchCtx {
  RowStore* rs = nullptr;
  int64_t n_rows = 0, nq = 0, nq_pad = 0;
  int k = 0, mode = 0, cap = 4096;
  bool tensor = false, l2_tensor = false, rescoring = false;
  Bf16ScanPlan plan{};
  alignas(64) unsigned char tmap_q[128];
  alignas(64) unsigned char tmap_db[128];
  int64_t kernels = 0, waves = 0;  // filled by enqueue_search
};

static int replan(sss_index* ix, SearchCtx& c) {
  if (!c.tensor) return 0;
  if (plan_scan_bf16(ix->d_pad, c.nq_pad, ix->num_sms, 8, &c.plan, ix->rec_boost,
                     ix->d + (ix->metric == SSS_METRIC_L2 ? 2 : 0), ix->tune.variant))
    return 1;
  c.plan.fp8 = ix->binary;
  return ix->ws.ensure_records(2 * c.plan.n_regions, c.plan.rec_cap);
}

// Enqueue one whole search on `st`: query staging, bootstrap thresholds, scan + refine waves, emit.  No allocation and
// no synchronisation in here (it is what gets captured into a graph); `profile` adds CUDA events around the scans.
static int enqueue_search(sss_index* ix, SearchCtx& c, const float* q_in, float* Ddev, int64_t* Idev, int* status_out,
                          bool safe, bool profile, cudaStream_t st, bool cautious = false) {
  Workspace& ws = ix->ws;
  RowStore& rs = *c.rs;
  const int64_t n_rows = c.n_rows;
  const bool tensor = c.tensor;
  const Bf16ScanPlan& plan = c.plan;
  c.kernels = c.waves = 0;
  SelectState state = ws.state();
  state.cap = c.cap;
  const int slack = !tensor ? 0 : c.mode == SSS_MODE_EXACT ? 1 : 2;
  if (ix->binary) {
    if (launch_prep_binary((const uint8_t*)q_in, c.nq, c.nq_pad, ix->b_nbytes, ix->d_pad * 2,
                           tensor ? (uint8_t*)ws.q_bf16 : nullptr, tensor ? nullptr : (uint8_t*)ws.q_keep, ix->d * 4, state, st))
      return 1;
  } else if (launch_prep_queries(q_in, c.nq, c.nq_pad, ix->d, ix->d_pad, tensor ? ws.q_bf16 : nullptr, slack, rs.maxnorm2,
                                 state, st, c.l2_tensor ? 1 : 0, ws.q_keep)) {
    return 1;
  }
  SSS_CUDA_OK(cudaMemsetAsync(ws.flags + 1, 0, sizeof(int), st));
  c.kernels += 1;
  RefineArgs ra;
  ra.nq = c.nq;
  ra.k = c.k;
  ra.reduce_max = ix->reduce == SSS_REDUCE_MAX;
  ra.row_seg = ix->row_seg;
  ra.db_f32 = rs.f32;
  ra.q_f32 = ws.q_keep;
  ra.d = ix->d;
  ra.metric = ix->metric;
  ra.debug = nullptr;
  if (profile && ix->dbg) {
    SSS_CUDA_OK(cudaMemsetAsync(ix->dbg, 0, 12 * sizeof(unsigned long long), st));
    ra.debug = ix->dbg;
  }
  // Bootstrap: one tensor-core pass in chunk-max mode over the first rows (128K; for smaller indexes the largest
  // power of two within half of the rows) gives every query a valid threshold at once and replaces the short first
  // waves (and, with re-scoring, the fp32 first wave and the per-wave re-scoring: lazy mode needs tensor waves only).
  int64_t boot_rows = 131072;
  while (boot_rows > 4096 && boot_rows * 2 > n_rows) boot_rows >>= 1;
  const int n_boot_chunks = (int)(boot_rows / 32);
  const bool grouped = ix->reduce == SSS_REDUCE_MAX;
  const int chunk_gap = grouped ? (int)((ix->max_seg_len + 30) / 32) + 1 : 1;
  const bool bootstrap = tensor && (plan.ts || plan.two_cta || plan.kloop) && !safe && n_rows >= 2 * boot_rows &&
                         !ix->tune.no_bootstrap && (int64_t)(c.k - 1) * chunk_gap + 1 <= n_boot_chunks / 4 &&
                         ws.cmax_elems >= (size_t)n_boot_chunks * (size_t)c.nq_pad;
  if (bootstrap) {
    if (launch_scan_bf16(plan, c.tmap_q, c.tmap_db, ws.q_bf16, 0, boot_rows, state, ws.rec, ws.rec_cnt, ws.flags + 1,
                         ws.cmax, st))
      return 1;
    if (launch_bootstrap_thr(ws.cmax, n_boot_chunks, c.nq, c.nq_pad, c.k, chunk_gap, 2.0f, state, st)) return 1;
    c.kernels += 2;
  }
  const std::vector<int64_t> ends =
      make_waves(n_rows, c.cap, c.k, safe, grouped, bootstrap ? boot_rows : 0, ix->tune, cautious && bootstrap,
                 /*quad=*/ix->binary && !tensor && !cautious);  // popcount scan: integer distances, strict filter — x4 waves
  int64_t begin = 0;
  uint32_t wave_id = 0;
  for (int64_t end : ends) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (profile) {
      while (ix->ev.size() < ix->ev_used + 2) {
        cudaEvent_t e;
        SSS_CUDA_OK(cudaEventCreate(&e));
        ix->ev.push_back(e);
      }
      e0 = ix->ev[ix->ev_used++];
      e1 = ix->ev[ix->ev_used++];
    }
    // The first wave has no threshold: every score is a candidate.  With re-scoring it is cheaper to get those
    // scores exactly from the fp32 scan than to rescore all of them after a tensor-core pass.
    const bool wave_tensor = tensor && !(c.rescoring && begin == 0 && !bootstrap);
    const uint32_t w = ++wave_id;
    const int buf = (int)(w & 1u);
    HitRecord* rec_buf = ws.rec + (tensor ? (size_t)buf * plan.n_regions * plan.rec_cap : 0);
    uint32_t* cnt_buf = ws.rec_cnt + (tensor ? (size_t)buf * plan.n_regions : 0);
    ra.rescore = c.rescoring && wave_tensor;
    ra.wave = w;
    ra.rec = wave_tensor ? rec_buf : nullptr;
    ra.rec_cnt = cnt_buf;
    ra.rec_nsub = tensor ? plan.rec_nsub : 0;
    ra.rec_cap = tensor ? plan.rec_cap : kRecSubCap;
    ra.l2_tensor = c.l2_tensor ? 1 : 0;
    ra.lazy = (c.rescoring && !safe && c.k <= 256 && !ix->tune.no_lazy) ? 1 : 0;
    ra.final = end == ends.back() ? 1 : 0;
    ra.row_limit = n_rows;
    if (e0) SSS_CUDA_OK(cudaEventRecord(e0, st));
    if (wave_tensor) {
      if (launch_scan_bf16(plan, c.tmap_q, c.tmap_db, ws.q_bf16, begin, end, state, rec_buf, cnt_buf, ws.flags + 1,
                           nullptr, st))
        return 1;
    } else if (ix->binary) {
      if (launch_scan_hamming((const uint8_t*)rs.f32, ix->d * 4, begin, end, (const uint8_t*)ws.q_keep, c.nq, state, st)) return 1;
    } else {
      if (launch_scan_fp32(rs.f32, ix->d, ix->metric, begin, end, ws.q_keep, c.nq, state, st)) return 1;
    }
    if (e1) SSS_CUDA_OK(cudaEventRecord(e1, st));
    if (launch_refine(ra, state, ix->num_sms, st)) return 1;
    c.kernels += 1 + (c.k <= 256 ? 2 : 1);
    c.waves += 1;
    begin = end;
  }
  if (ix->binary) {
    if (launch_emit_hamming(state, c.nq, c.k, tensor ? ix->b_nbits : 0, ix->id_offset, (int32_t*)Ddev, Idev, st)) return 1;
  } else if (launch_emit(state, c.nq, c.k, ix->metric, ix->id_offset, Ddev, Idev, status_out, st)) {
    return 1;
  }
  c.kernels += 1;
  return 0;
}

static int find_graph_nodes(SearchGraph& g, bool binary) {
  const void* prep_fn = binary ? prep_binary_kernel_addr() : prep_queries_kernel_addr();
  const void* emit_fn = binary ? emit_hamming_kernel_addr() : emit_kernel_addr();
  size_t n = 0;
  SSS_CUDA_OK(cudaGraphGetNodes(g.graph, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  SSS_CUDA_OK(cudaGraphGetNodes(g.graph, nodes.data(), &n));
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType t;
    SSS_CUDA_OK(cudaGraphNodeGetType(nodes[i], &t));
    if (t != cudaGraphNodeTypeKernel) continue;
    cudaKernelNodeParams kp;
    SSS_CUDA_OK(cudaGraphKernelNodeGetParams(nodes[i], &kp));
    if (kp.func == prep_fn) g.prep = nodes[i];
    if (kp.func == emit_fn) g.emit = nodes[i];
  }
  SSS_REQUIRE(g.prep != nullptr && g.emit != nullptr, "captured search graph lacks its staging / emit nodes");
  return 0;
}

// re-bind pointer arguments of one kernel node: (argument index, new value) pairs
static int rebind(SearchGraph& g, cudaGraphNode_t node, int n_args, int i0, const void* v0, int i1 = -1,
                  const void* v1 = nullptr, int i2 = -1, const void* v2 = nullptr) {
  cudaKernelNodeParams kp;
  SSS_CUDA_OK(cudaGraphKernelNodeGetParams(node, &kp));
  SSS_REQUIRE(kp.kernelParams != nullptr && n_args <= 16, "captured kernel node exposes no parameter array");
  void* args[16];
  for (int i = 0; i < n_args; ++i) args[i] = kp.kernelParams[i];
  args[i0] = (void*)&v0;
  if (i1 >= 0) args[i1] = (void*)&v1;
  if (i2 >= 0) args[i2] = (void*)&v2;
  kp.kernelParams = args;
  kp.extra = nullptr;
  SSS_CUDA_OK(cudaGraphExecKernelNodeSetParams(g.exec, node, &kp));
  return 0;
}

static SearchGraph* capture_graph(sss_index* ix, SearchCtx& c, const float* q_in, float* Ddev, int64_t* Idev,
                                  int* status_out) {
  if (!ix->cap_stream && cudaStreamCreateWithFlags(&ix->cap_stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  SearchGraph g;
  g.nq = c.nq;
  g.k = c.k;
  g.mode = c.mode;
  g.epoch = ix->epoch;
  g.ws_gen = ix->ws.generation;
  if (cudaStreamBeginCapture(ix->cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  const int rc = enqueue_search(ix, c, q_in, Ddev, Idev, status_out, false, false, ix->cap_stream);
  cudaError_t e = cudaStreamEndCapture(ix->cap_stream, &g.graph);
  if (rc != 0 || e != cudaSuccess || g.graph == nullptr) {
    cudaGetLastError();
    g.destroy();
    return nullptr;
  }
  if (cudaGraphInstantiate(&g.exec, g.graph, 0) != cudaSuccess || find_graph_nodes(g, ix->binary)) {
    cudaGetLastError();
    g.destroy();
    return nullptr;
  }
  g.kernels = c.kernels;
  g.waves = c.waves;
  if (ix->graphs.size() >= 6) {  // a handful of (nq, k, mode) shapes per index: drop the least recently used
    size_t lru = 0;
    for (size_t i = 1; i < ix->graphs.size(); ++i)
      if (ix->graphs[i].last_use < ix->graphs[lru].last_use) lru = i;
    ix->graphs[lru].destroy();
    ix->graphs.erase(ix->graphs.begin() + lru);
  }
  ix->graphs.push_back(g);
  return &ix->graphs.back();
}

static int search_batch(sss_index* ix, const BatchArgs& b, cudaStream_t st) {
  const bool use_sums = ix->reduce == SSS_REDUCE_SUM;
  SearchCtx c;
  c.rs = use_sums ? &ix->sums : &ix->rows;
  c.n_rows = c.rs->n;
  c.nq = b.nq;
  c.k = b.k;
  c.mode = b.mode;
  if (c.mode != SSS_MODE_FP32 && !ix->tensor_ok) c.mode = SSS_MODE_FP32;  // no tensor path for d_pad > 4096: the fp32 scan IS the exact result
  c.tensor = c.mode != SSS_MODE_FP32 && c.n_rows > 0;
  c.l2_tensor = c.tensor && ix->metric == SSS_METRIC_L2;
  c.rescoring = c.mode != SSS_MODE_FP32;
  if (ix->binary) {
    // codes of up to 256 bits: the +-1 fp8 tensor-core scan (at 100M x 256 bit: 63 K queries/s = 3.2 PFLOP/s fp8 at
    // nq = 1000, and a 6.3 TB/s stream of the one-byte-per-bit rows at nq <= 128: profiles/r02_binary.md).  Up to
    // kHammingSmallNq queries, and for longer codes, the popcount scan over the PACKED codes: an eighth of the bytes.
    c.tensor = ix->tensor_ok && c.n_rows > 0 && b.nq > kHammingSmallNq;
    c.rescoring = false;
    c.l2_tensor = false;
    c.mode = c.tensor ? SSS_MODE_BF16 : SSS_MODE_FP32;  // (graph key: the two scans are different graphs)
  }
  ix->stat_variant = 0;
  ix->stat_graph = 0;
  SSS_REQUIRE(b.k <= c.cap / 2, "k too large (max 2048)");
  c.nq_pad = (b.nq + 127) / 128 * 128;
  Workspace& ws = ix->ws;
  if (ws.ensure(c.nq_pad, c.cap, ix->d, ix->d_pad, b.out_on_device ? 0 : b.nq * b.k)) return 1;
  if (replan(ix, c)) return 1;
  if (c.tensor) {
    ix->stat_variant = c.plan.kloop ? 4 : c.plan.two_cta ? 3 : c.plan.ts ? 2 : 1;
    if (ws.ensure_cmax((size_t)4096 * (size_t)c.nq_pad)) return 1;
    if (make_tensor_map_bf16_2d(c.tmap_q, ws.q_bf16, (uint64_t)c.nq_pad, (uint64_t)ix->d_pad, 128)) return 1;
    if (make_tensor_map_bf16_2d(c.tmap_db, c.rs->bf16, (uint64_t)c.n_rows, (uint64_t)ix->d_pad, 128)) return 1;
  }
  if (ix->profile && !ix->dbg && dev_alloc(&ix->dbg, 12)) return 1;
  const float* qdev = b.q;
  if (!b.q_on_device) {
    SSS_CUDA_OK(cudaMemcpyAsync(ws.q_f32, b.q, (size_t)b.nq * (size_t)ix->q_row_bytes(), cudaMemcpyHostToDevice, st));
    qdev = ws.q_f32;
  }
  float* Ddev = b.out_on_device ? b.D : ws.out_D;
  int64_t* Idev = b.out_on_device ? b.I : ws.out_I;

  // Attempts: the normal schedule (replayed from a captured graph when there is one); if ONLY a record sub-region
  // overflowed (a hot spot of one query inside one CTA's share of a wave), the same schedule again with 4x the records
  // per sub-region (the index remembers that); anything else, or a second overflow, falls back to the safe schedule
  // of 2048-row waves.
  bool safe = false, cautious = false;
  for (int attempt = 0; attempt < 4; ++attempt) {
    if (c.tensor && c.plan.rec_cap != (c.plan.kloop ? 4 : 1) * kRecSubCap * ix->rec_boost && replan(ix, c)) return 1;
    bool launched = false;
    if (!safe && !cautious && !ix->profile && !ix->tune.no_graph) {
      SearchGraph* g = nullptr;
      for (auto& cand : ix->graphs)
        if (cand.nq == c.nq && cand.k == c.k && cand.mode == c.mode && cand.epoch == ix->epoch &&
            cand.ws_gen == ws.generation)
          g = &cand;
      if (!g) g = capture_graph(ix, c, qdev, Ddev, Idev, b.status_out);
      if (g && rebind(*g, g->prep, ix->binary ? 8 : 11, 0, qdev) == 0 &&
          rebind(*g, g->emit, ix->binary ? 7 : 8, 5, Ddev, 6, Idev, ix->binary ? -1 : 7, b.status_out) == 0 &&
          cudaGraphLaunch(g->exec, st) == cudaSuccess) {
        g->last_use = ++ix->graph_clock;
        c.kernels = g->kernels;
        c.waves = g->waves;
        ix->stat_graph = 1;
        launched = true;
      } else {
        cudaGetLastError();  // graphs are an optimisation: anything unexpected falls back to plain launches
      }
    }
    if (!launched && enqueue_search(ix, c, qdev, Ddev, Idev, b.status_out, safe, ix->profile, st, cautious)) return 1;
    ix->stat_kernels += c.kernels;
    ix->stat_waves += c.waves;
    // asynchronous form: the caller reads the status word when it suits it (a profiled search stays synchronous: its
    // CUDA events are read below)
    if (b.status_out != nullptr && !ix->profile) return 0;
    SSS_CUDA_OK(cudaMemcpyAsync(ws.host_flags, ws.flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (!b.out_on_device) {
      SSS_CUDA_OK(cudaMemcpyAsync(b.D, Ddev, (size_t)b.nq * b.k * sizeof(float), cudaMemcpyDeviceToHost, st));
      SSS_CUDA_OK(cudaMemcpyAsync(b.I, Idev, (size_t)b.nq * b.k * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    }
    if (ix->profile && ix->dbg)
      SSS_CUDA_OK(cudaMemcpyAsync(ix->dbg_host, ix->dbg, sizeof(ix->dbg_host), cudaMemcpyDeviceToHost, st));
    SSS_CUDA_OK(cudaStreamSynchronize(st));
    if (ix->profile) {
      for (size_t i = 0; i + 1 < ix->ev_used; i += 2) {
        float ms = 0.0f;
        SSS_CUDA_OK(cudaEventElapsedTime(&ms, ix->ev[i], ix->ev[i + 1]));
        ix->scan_us += (double)ms * 1000.0;
        ix->scan_launches += 1;
      }
      ix->ev_used = 0;
    }
    const int f_over = ws.host_flags[0], f_dog = ws.host_flags[1];
    SSS_REQUIRE(f_dog == 0, "tensor-core scan watchdog fired (barrier code " + std::to_string(f_dog) + ")");
    if (f_over == 0) return 0;
    // a candidate list or record region overflowed (adversarial score order): redo with waves that
    // cannot overflow by construction
    SSS_REQUIRE(!safe, "candidate overflow persisted in the safe wave schedule (internal error)");
    ix->stat_reruns += 1;
    ix->stat_overflow_reason = f_over;
    if (c.tensor && f_over == 1 && ix->rec_boost == 1) {
      ix->rec_boost = 4;
      ix->epoch += 1;
      ix->drop_graphs();
    } else if (c.tensor && !cautious && !safe) {
      cautious = true;   // same machinery, waves growing x1.4 from half the bootstrap sample
    } else {
      cautious = false;
      safe = true;
    }
  }
  return 0;
}

static int search_all(sss_index* ix, const float* q, int64_t nq, int k, int mode, int q_on_device, float* D, int64_t* I,
                      int out_on_device, void* stream, const char* who, int* status_out = nullptr) {
  SSS_REQUIRE(ix != nullptr, std::string(who) + ": NULL index");
  SSS_REQUIRE(k >= 1, std::string(who) + ": k must be >= 1");
  SSS_REQUIRE(nq >= 0, std::string(who) + ": negative nq");
  SSS_REQUIRE(mode >= 0 && mode <= 2, std::string(who) + ": unknown mode");
  if (nq == 0) return 0;
  SSS_REQUIRE(q && D && I, std::string(who) + ": NULL buffer");
  DeviceGuard g(ix->device);
  SSS_REQUIRE(g.ok, std::string(who) + ": cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  ix->stat_kernels = ix->stat_waves = ix->stat_reruns = 0;
  ix->scan_us = 0.0;
  ix->scan_launches = 0;
  const int64_t QB = 2048;  // queries per pass: <= 16 resident m-tile groups, grid_x >= 37
  for (int64_t q0 = 0; q0 < nq; q0 += QB) {
    BatchArgs b;
    b.nq = std::min(QB, nq - q0);
    b.q = (const float*)((const char*)q + q0 * ix->q_row_bytes());
    b.k = k;
    b.mode = mode;
    b.q_on_device = q_on_device != 0;
    b.out_on_device = out_on_device != 0;
    b.D = D + q0 * k;
    b.I = I + q0 * k;
    b.status_out = status_out;
    if (search_batch(ix, b, st)) return 1;
  }
  return 0;
}

}  // namespace sss

extern "C" int sss_index_search(sss_index_t* ix, const float* q, int64_t nq, int k, int mode, int q_on_device,
                                float* D, int64_t* I, int out_on_device, void* stream) {
  return search_all(ix, q, nq, k, mode, q_on_device, D, I, out_on_device, stream, "sss_index_search");
}

// Bytes of one rank's packed candidate block [ids int64 nq*k | scores fp32 nq*k], padded to 16 bytes: the unit of
// the sharded search's single all-gather.
// + 16 bytes of trailer: [int32 status | 12 bytes pad]
extern "C" int64_t sss_packed_bytes(int64_t nq, int k) { return (nq * k * 12 + 15) / 16 * 16 + 16; }

extern "C" int sss_index_search_packed(sss_index_t* ix, const float* q, int64_t nq, int k, int mode, int q_on_device,
                                       void* packed, int async, void* stream) {
  SSS_REQUIRE(packed != nullptr, "sss_index_search_packed: NULL buffer");
  SSS_REQUIRE(nq >= 0 && k >= 1, "sss_index_search_packed: bad nq / k");
  int64_t* I = (int64_t*)packed;
  float* D = (float*)((char*)packed + (size_t)nq * k * 8);
  int* status = (int*)((char*)packed + sss_packed_bytes(nq, k) - 16);
  {
    DeviceGuard g(ix ? ix->device : 0);
    SSS_CUDA_OK(cudaMemsetAsync(status, 0, 16, (cudaStream_t)stream));
  }
  return search_all(ix, q, nq, k, mode, q_on_device, D, I, 1, stream, "sss_index_search_packed", async ? status : nullptr);
}

extern "C" int sss_normalize(const float* in, float* out, int64_t n, int d, int norm_mode, int on_device, int device,
                             void* stream) {
  SSS_REQUIRE(in && out, "sss_normalize: NULL buffer");
  SSS_REQUIRE(n >= 0 && d >= 1, "sss_normalize: bad shape");
  SSS_REQUIRE(norm_mode >= 0 && norm_mode <= 3, "sss_normalize: unknown norm_mode");
  if (n == 0) return 0;
  DeviceGuard g(device);
  SSS_REQUIRE(g.ok, "sss_normalize: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  if (on_device) return launch_normalize(in, out, n, d, norm_mode, st);
  float* tmp = nullptr;
  if (dev_alloc(&tmp, (size_t)n * d)) return 1;
  int rc = 0;
  if (cudaMemcpyAsync(tmp, in, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = 1;
  if (!rc) rc = launch_normalize(tmp, tmp, n, d, norm_mode, st);
  if (!rc && cudaMemcpyAsync(out, tmp, (size_t)n * d * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = 1;
  if (cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
  cudaFree(tmp);
  if (rc && g_err.empty()) set_error("sss_normalize: CUDA copy failed");
  return rc;
}

extern "C" int sss_gather_rows(const float* table, int64_t n_rows, int d, const int64_t* ids, int64_t n, float* out,
                               int device, void* stream) {
  SSS_REQUIRE(n_rows >= 0 && d >= 1 && n >= 0, "sss_gather_rows: bad shape");
  if (n == 0) return 0;
  SSS_REQUIRE(table && out && ids, "sss_gather_rows: NULL buffer");
  DeviceGuard g(device);
  SSS_REQUIRE(g.ok, "sss_gather_rows: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  int* bad = nullptr;
  if (dev_alloc(&bad, 1)) return 1;
  int host_bad = 0, rc = 0;
  if (cudaMemsetAsync(bad, 0, sizeof(int), st) != cudaSuccess) rc = 1;
  if (!rc) rc = launch_gather_rows(table, n_rows, d, ids, n, out, bad, st);
  if (!rc && cudaMemcpyAsync(&host_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = 1;
  if (cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
  cudaFree(bad);
  if (rc && g_err.empty()) set_error("sss_gather_rows: CUDA call failed");
  SSS_REQUIRE(rc != 0 || host_bad == 0, "index out of range in self");   // torch.nn.Embedding's message
  return rc;
}

extern "C" int sss_topk_merge(const float* cand_D, const int64_t* cand_I, int n_shards, int64_t nq, int k, int metric,
                              float* D, int64_t* I, int device, void* stream) {
  SSS_REQUIRE(cand_D && cand_I && D && I, "sss_topk_merge: NULL buffer");
  SSS_REQUIRE(n_shards >= 1 && k >= 1 && nq >= 0, "sss_topk_merge: bad shape");
  DeviceGuard g(device);
  SSS_REQUIRE(g.ok, "sss_topk_merge: cudaSetDevice failed");
  return launch_topk_merge(cand_D, cand_I, nq * k, nq * k, n_shards, nq, k, metric, D, I, (cudaStream_t)stream);
}

extern "C" int sss_topk_merge_packed(const void* gathered, int n_shards, int64_t nq, int k, int metric, float* D,
                                     int64_t* I, int* status_out, int device, void* stream) {
  SSS_REQUIRE(gathered && D && I, "sss_topk_merge_packed: NULL buffer");
  SSS_REQUIRE(n_shards >= 1 && k >= 1 && nq >= 0, "sss_topk_merge_packed: bad shape");
  DeviceGuard g(device);
  SSS_REQUIRE(g.ok, "sss_topk_merge_packed: cudaSetDevice failed");
  const int64_t block = sss_packed_bytes(nq, k);
  const int64_t* cI = (const int64_t*)gathered;
  const float* cD = (const float*)((const char*)gathered + (size_t)nq * k * 8);
  const int* st_in = (const int*)((const char*)gathered + block - 16);
  return launch_topk_merge(cD, cI, block / 4, block / 8, n_shards, nq, k, metric, D, I, (cudaStream_t)stream, st_in,
                           block / 4, status_out);
}

// ---------------------------------------------------------------------------------------------------
// binary index: a float index in its binary flavour (packed codes in the fp32 slot of the row store, +-1 E4M3
// rows in the bf16 slot), driven by the same wave schedule, refine and graph replay
// ---------------------------------------------------------------------------------------------------
namespace sss {
__global__ void repack_codes_kernel(const uint8_t* __restrict__ in, int nbytes, uint8_t* __restrict__ out, int pitch,
                                    int64_t n) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * pitch) return;
  int64_t r = t / pitch;
  int b = (int)(t % pitch);
  out[t] = b < nbytes ? in[r * nbytes + b] : (uint8_t)0;
}
static int launch_repack(const uint8_t* in, int nbytes, uint8_t* out, int pitch, int64_t n, cudaStream_t st) {
  int64_t total = n * pitch;
  if (total <= 0) return 0;
  repack_codes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, nbytes, out, pitch, n);
  SSS_CUDA_OK(cudaGetLastError());
  return 0;
}
}  // namespace sss

struct sss_binary_index {
  sss_index core;
};

extern "C" int sss_binary_create(sss_binary_index_t** out, int device, int nbits, int64_t id_offset) {
  SSS_REQUIRE(out != nullptr, "sss_binary_create: out is NULL");
  SSS_REQUIRE(nbits >= 8 && nbits % 8 == 0 && nbits <= 512, "sss_binary_create: nbits must be a multiple of 8, <= 512");
  int ndev = 0;
  SSS_CUDA_OK(cudaGetDeviceCount(&ndev));
  SSS_REQUIRE(device >= 0 && device < ndev, "sss_binary_create: no such CUDA device");
  cudaDeviceProp prop;
  SSS_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  SSS_REQUIRE(prop.major == 10, "libsss_b200 is built for sm_100a only; device is sm_" +
                                    std::to_string(prop.major * 10 + prop.minor));
  sss_binary_index* bx = new sss_binary_index();
  sss_index& ix = bx->core;
  ix.device = device;
  ix.binary = true;
  ix.b_nbits = nbits;
  ix.b_nbytes = nbits / 8;
  ix.d = (ix.b_nbytes + 3) / 4;                  // packed pitch in 4-byte words: the fp32 slot holds the codes
  ix.d_pad = (nbits + 127) / 128 * 128 / 2;      // fp8 row bytes / 2: the geometry of a bf16 row of that many bytes
  ix.metric = SSS_METRIC_IP;
  ix.id_offset = id_offset;
  ix.num_sms = prop.multiProcessorCount;
  ix.tensor_ok = ix.d_pad <= 128;                // the pair / TS kernels: codes of up to 256 bits (the reference: 250)
  ix.tune = Tuning::from_env();
  *out = bx;
  return 0;
}

extern "C" int sss_binary_destroy(sss_binary_index_t* bx) {
  if (!bx) return 0;
  sss_index& ix = bx->core;
  DeviceGuard g(ix.device);
  if (g.ok) {
    cudaDeviceSynchronize();
    ix.drop_graphs();
    ix.rows.release();
    dev_free(ix.dbg);
    ix.ws.release();
    for (cudaEvent_t e : ix.ev) cudaEventDestroy(e);
    if (ix.cap_stream) cudaStreamDestroy(ix.cap_stream);
  }
  delete bx;
  return g.ok ? 0 : 1;
}

extern "C" int64_t sss_binary_ntotal(const sss_binary_index_t* bx) { return bx ? bx->core.rows.n : 0; }
extern "C" int64_t sss_binary_stat(const sss_binary_index_t* bx, int what) { return bx ? sss_index_stat(&bx->core, what) : -1; }
extern "C" int sss_binary_set_profiling(sss_binary_index_t* bx, int on) {
  SSS_REQUIRE(bx != nullptr, "sss_binary_set_profiling: NULL index");
  bx->core.profile = on != 0;
  return 0;
}

extern "C" int sss_binary_add(sss_binary_index_t* bx, const uint8_t* codes, int64_t n, int on_device, void* stream) {
  SSS_REQUIRE(bx != nullptr, "sss_binary_add: NULL index");
  SSS_REQUIRE(n >= 0, "sss_binary_add: negative row count");
  if (n == 0) return 0;
  SSS_REQUIRE(codes != nullptr, "sss_binary_add: NULL codes");
  sss_index& ix = bx->core;
  SSS_REQUIRE(ix.rows.n + n < 0xFFFFFFF0ll, "sss_binary_add: more than 2^32 rows in one shard");
  DeviceGuard g(ix.device);
  SSS_REQUIRE(g.ok, "sss_binary_add: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  ix.epoch += 1;
  ix.drop_graphs();
  if (ix.rows.ensure(ix.rows.n + n, ix.d, ix.d_pad, ix.tensor_ok, st)) return 1;
  const uint8_t* src = codes;
  uint8_t* staged = nullptr;
  if (!on_device) {
    if (dev_alloc(&staged, (size_t)n * ix.b_nbytes)) return 1;
    if (cudaMemcpyAsync(staged, codes, (size_t)n * ix.b_nbytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      cudaFree(staged);
      set_error("sss_binary_add: host to device copy failed");
      return 1;
    }
    src = staged;
  }
  const int pitch = ix.d * 4;
  int rc = launch_repack(src, ix.b_nbytes, (uint8_t*)ix.rows.f32 + (size_t)ix.rows.n * pitch, pitch, n, st);
  if (!rc && ix.tensor_ok)
    rc = launch_expand_codes_fp8(src, ix.b_nbytes, n, ix.d_pad * 2, (uint8_t*)ix.rows.bf16 + (size_t)ix.rows.n * ix.d_pad * 2, st);
  if (staged) {
    cudaStreamSynchronize(st);
    cudaFree(staged);
  }
  if (rc) return rc;
  ix.rows.n += n;
  return 0;
}

extern "C" int sss_binary_search(sss_binary_index_t* bx, const uint8_t* q, int64_t nq, int k, int q_on_device,
                                 int32_t* D, int64_t* I, int out_on_device, void* stream) {
  SSS_REQUIRE(bx != nullptr, "sss_binary_search: NULL index");
  // (integer distances travel through the float index's driver: same 4-byte slots, written by the Hamming emit)
  return search_all(&bx->core, (const float*)q, nq, k, SSS_MODE_EXACT, q_on_device, (float*)D, I, out_on_device, stream,
                    "sss_binary_search");
}

extern "C" int sss_pack_sign_bits(const float* x, uint8_t* codes, int64_t n, int nbits_in, int on_device, int device,
                                  void* stream) {
  SSS_REQUIRE(x && codes, "sss_pack_sign_bits: NULL buffer");
  SSS_REQUIRE(n >= 0 && nbits_in >= 1, "sss_pack_sign_bits: bad shape");
  if (n == 0) return 0;
  DeviceGuard g(device);
  SSS_REQUIRE(g.ok, "sss_pack_sign_bits: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  if (on_device) return launch_pack_sign_bits(x, codes, n, nbits_in, st);
  const int nbytes = (nbits_in + 7) / 8;
  float* dx = nullptr;
  uint8_t* dc = nullptr;
  if (dev_alloc(&dx, (size_t)n * nbits_in) || dev_alloc(&dc, (size_t)n * nbytes)) return 1;
  int rc = 0;
  if (cudaMemcpyAsync(dx, x, (size_t)n * nbits_in * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = 1;
  if (!rc) rc = launch_pack_sign_bits(dx, dc, n, nbits_in, st);
  if (!rc && cudaMemcpyAsync(codes, dc, (size_t)n * nbytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = 1;
  if (cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
  cudaFree(dx);
  cudaFree(dc);
  if (rc && g_err.empty()) set_error("sss_pack_sign_bits: CUDA failure");
  return rc;
}
