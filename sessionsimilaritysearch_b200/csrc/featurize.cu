// featurize.cu — batched HOST featuriser (no device code): flattened session actions -> the integer arrays the
// encoder reads from a batch of session graphs.  It replaces the per-session Python of the reference's
// sequence_to_graph (util_amazon_filtered.py:98-230 with helpers :7-22, :75-83) followed by PyG's
// Batch.from_data_list (test_amazon_filterd.py:485-488) for everything the encoder consumes: node order, positions,
// counts, the three edge lists with batch-global indices, transition multiplicities and the last-click mask.
// Text is not touched here: a node carries the KEY of its text (the caller's id of the query string, or the item
// id), and the features are gathered from a cache on the device.
//
// Layout rules restated from the reference (same as sessionsimilaritysearch_b200/sessions.py, the Python mirror):
//   query nodes    node 0 = empty-string root (pos = len), then one per search at position i (pos = len - (i + 1))
//   product nodes  the session's distinct items in the order the caller gives (the reference: list(set(ids)));
//                  cnt = occurrences; pos = len - j for every occurrence j, grouped by product;
//                  an item-less session gets one node (item 0, cnt 1, pos 0)
//   q -> p edges   one per item event, from the latest search node (root before the first search); multi-edges kept
//   p -> p edges   consecutive item transitions, de-duplicated in first-appearance order with their multiplicity;
//                  self transitions kept
//   last click     destination of the last transition (node 0 when there is none)
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sss_b200.h"
#include "common.cuh"

namespace {

// A persistent worker pool: creating a thread costs ~2 ms in this kind of sandbox (measured: eight std::thread
// constructors = 16 ms, more than the work they were given), so the workers are started once and parked on a
// condition variable.  One parallel region at a time; a caller that finds the pool busy runs its region inline.
class WorkerPool {
 public:
  static WorkerPool& get() {
    static WorkerPool* p = new WorkerPool();  // never destroyed: the workers are detached and outlive static teardown
    return *p;
  }
  // fn(chunk) for chunk in [0, n_chunks), on up to max_threads threads including the caller
  void run(int n_chunks, int max_threads, const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> region(region_m_, std::try_to_lock);
    if (n_chunks <= 1 || max_threads <= 1 || !region.owns_lock()) {
      for (int c = 0; c < n_chunks; ++c) fn(c);
      return;
    }
    ensure_workers(std::min(max_threads, n_chunks) - 1);
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &fn;
      next_.store(0);
      n_chunks_ = n_chunks;
      active_ = std::min<int>((int)workers_, std::min(max_threads, n_chunks) - 1);
      running_ = active_;
      ++generation_;
    }
    cv_.notify_all();
    for (int c = next_.fetch_add(1); c < n_chunks; c = next_.fetch_add(1)) fn(c);
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return running_ == 0; });
    job_ = nullptr;
  }

 private:
  void ensure_workers(int want) {
    std::lock_guard<std::mutex> lk(m_);
    while ((int)workers_ < want) {
      const int id = (int)workers_++;
      std::thread([this, id] { loop(id); }).detach();
    }
  }
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* job = nullptr;
      int n = 0;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (id >= active_) continue;   // not part of this region
        job = job_;
        n = n_chunks_;
      }
      for (int c = next_.fetch_add(1); c < n; c = next_.fetch_add(1)) (*job)(c);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--running_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::mutex region_m_, m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* job_ = nullptr;
  std::atomic<int> next_{0};
  int n_chunks_ = 0, active_ = 0, running_ = 0;
  size_t workers_ = 0;
  uint64_t generation_ = 0;
};

struct Counts {
  int64_t nq = 0, np = 0, ne = 0, eqp = 0, epp = 0;
};

inline int slot_of(const int64_t* uniq, int n_uniq, int64_t item) {
  for (int i = 0; i < n_uniq; ++i)
    if (uniq[i] == item) return i;
  return -1;
}

// distinct consecutive transitions of one session, first-appearance order; returns their number
inline int transitions(const sss_flat_sessions_t* s, int64_t sess, std::vector<int>& chain, std::vector<int>& pa,
                       std::vector<int>& pb, std::vector<int>& pw, bool* bad) {
  const int64_t a0 = s->act_off[sess], a1 = s->act_off[sess + 1];
  const int64_t* uniq = s->uniq_items + s->uniq_off[sess];
  const int n_uniq = (int)(s->uniq_off[sess + 1] - s->uniq_off[sess]);
  chain.clear();
  pa.clear();
  pb.clear();
  pw.clear();
  for (int64_t a = a0; a < a1; ++a)
    if (!s->act_is_search[a]) {
      const int sl = slot_of(uniq, n_uniq, s->act_key[a]);
      if (sl < 0) {
        *bad = true;
        return 0;
      }
      chain.push_back(sl);
    }
  for (size_t i = 0; i + 1 < chain.size(); ++i) {
    const int x = chain[i], y = chain[i + 1];
    size_t j = 0;
    for (; j < pa.size(); ++j)
      if (pa[j] == x && pb[j] == y) break;
    if (j == pa.size()) {
      pa.push_back(x);
      pb.push_back(y);
      pw.push_back(1);
    } else {
      pw[j] += 1;
    }
  }
  return (int)pa.size();
}

int count_session(const sss_flat_sessions_t* s, int64_t sess, Counts* c, std::vector<int>& chain, std::vector<int>& pa,
                  std::vector<int>& pb, std::vector<int>& pw) {
  const int64_t a0 = s->act_off[sess], a1 = s->act_off[sess + 1];
  int64_t searches = 0, items = 0;
  for (int64_t a = a0; a < a1; ++a) (s->act_is_search[a] ? searches : items) += 1;
  const int64_t n_uniq = s->uniq_off[sess + 1] - s->uniq_off[sess];
  if ((items == 0) != (n_uniq == 0)) return 1;
  bool bad = false;
  c->nq = 1 + searches;
  c->np = n_uniq > 0 ? n_uniq : 1;
  c->ne = items > 0 ? items : 1;
  c->eqp = items;
  c->epp = transitions(s, sess, chain, pa, pb, pw, &bad);
  return bad ? 1 : 0;
}

}  // namespace

extern "C" int sss_featurize_sizes(const sss_flat_sessions_t* s, int64_t* n_query, int64_t* n_product,
                                   int64_t* n_expanded, int64_t* e_qp, int64_t* e_pp) {
  SSS_REQUIRE(s && s->n_sessions >= 0 && (s->n_sessions == 0 || (s->act_off && s->uniq_off)),
              "sss_featurize_sizes: bad argument");
  Counts tot;
  std::vector<int> chain, pa, pb, pw;
  for (int64_t i = 0; i < s->n_sessions; ++i) {
    Counts c;
    SSS_REQUIRE(count_session(s, i, &c, chain, pa, pb, pw) == 0,
                "sss_featurize: the distinct-item list of session " + std::to_string(i) +
                    " does not match its item events");
    tot.nq += c.nq; tot.np += c.np; tot.ne += c.ne; tot.eqp += c.eqp; tot.epp += c.epp;
  }
  if (n_query) *n_query = tot.nq;
  if (n_product) *n_product = tot.np;
  if (n_expanded) *n_expanded = tot.ne;
  if (e_qp) *e_qp = tot.eqp;
  if (e_pp) *e_pp = tot.epp;
  return 0;
}

// Shared body.  batch_size > 0: the sessions form consecutive encoder batches of that many sessions; the output
// arrays are the batches laid end to end, every node / graph index is LOCAL to its batch, and bounds[b * 5 ..] holds the
// (query, product, expanded, q->p, p->p) offsets of batch b in them (n_batches + 1 rows).
static int featurize_impl(const sss_flat_sessions_t* s, int64_t batch_size, int64_t root_query_key, sss_graph_arrays_t* out,
                          int64_t* bounds, int n_threads) {
  SSS_REQUIRE(s && out && s->n_sessions >= 0, "sss_featurize_batch: bad argument");
  const int64_t n = s->n_sessions;
  const int64_t bs = batch_size > 0 ? batch_size : (n > 0 ? n : 1);
  // worker threads for both passes (sessions are independent): at most 8, and at least 256 sessions each
  int nt = n_threads > 0 ? n_threads : std::min(8, (int)std::thread::hardware_concurrency());
  nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n / 256));
  // chunks of 512 sessions handed out dynamically to the pool
  const int64_t chunk = 512;
  const int n_chunks = (int)((n + chunk - 1) / chunk);
  auto parallel = [&](auto&& fn) {
    const std::function<void(int)> job = [&](int c) { fn((int64_t)c * chunk, std::min<int64_t>(n, ((int64_t)c + 1) * chunk)); };
    WorkerPool::get().run(n_chunks, nt, job);
  };
  // pass 1: per-session sizes (parallel) -> offsets (serial prefix)
  std::vector<Counts> off((size_t)n + 1);
  std::atomic<int64_t> bad_session{-1};
  parallel([&](int64_t lo, int64_t hi) {
    thread_local std::vector<int> chain, pa, pb, pw;
    for (int64_t i = lo; i < hi; ++i)
      if (count_session(s, i, &off[(size_t)i], chain, pa, pb, pw) != 0) bad_session.store(i);
  });
  SSS_REQUIRE(bad_session.load() < 0, "sss_featurize: the distinct-item list of session " +
                                           std::to_string(bad_session.load()) + " does not match its item events");
  {
    Counts run;
    for (int64_t i = 0; i < n; ++i) {
      const Counts c = off[(size_t)i];
      off[(size_t)i] = run;
      run.nq += c.nq; run.np += c.np; run.ne += c.ne; run.eqp += c.eqp; run.epp += c.epp;
    }
    off[(size_t)n] = run;
  }
  const Counts& tot = off[(size_t)n];
  SSS_REQUIRE(out->cap_query >= tot.nq && out->cap_product >= tot.np && out->cap_expanded >= tot.ne &&
                  out->cap_qp >= tot.eqp && out->cap_pp >= tot.epp,
              "sss_featurize_batch: output capacity too small (use sss_featurize_sizes)");
  out->n_query = tot.nq; out->n_product = tot.np; out->n_expanded = tot.ne; out->e_qp = tot.eqp; out->e_pp = tot.epp;
  if (bounds) {
    const int64_t nb = (n + bs - 1) / bs;
    for (int64_t b = 0; b <= nb; ++b) {
      const Counts& c = off[(size_t)std::min<int64_t>(n, b * bs)];
      bounds[b * 5 + 0] = c.nq; bounds[b * 5 + 1] = c.np; bounds[b * 5 + 2] = c.ne; bounds[b * 5 + 3] = c.eqp; bounds[b * 5 + 4] = c.epp;
    }
  }

  // pass 2: fill (sessions are independent: split them over threads)
  auto fill = [&](int64_t lo, int64_t hi) {
    thread_local std::vector<int> chain, pa, pb, pw;
    for (int64_t i = lo; i < hi; ++i) {
      const Counts& o = off[(size_t)i];              // output positions (over all batches)
      const int64_t i0 = i / bs * bs;                // first session of this session's batch
      const Counts& base = off[(size_t)i0];
      const int64_t v_nq = o.nq - base.nq, v_np = o.np - base.np;   // node index VALUES: local to the batch
      const int64_t gi = i - i0;                                     // graph index inside the batch
      const int64_t a0 = s->act_off[i], a1 = s->act_off[i + 1];
      const int64_t len = a1 - a0;
      const int64_t* uniq = s->uniq_items + s->uniq_off[i];
      const int n_uniq = (int)(s->uniq_off[i + 1] - s->uniq_off[i]);
      // query nodes + q -> p edges
      int64_t q = o.nq, eq = o.eqp;
      out->query_key[q] = root_query_key;
      out->query_pos[q] = len;
      out->query_batch[q] = gi;
      ++q;
      int64_t cur = 0;
      for (int64_t a = a0; a < a1; ++a) {
        if (s->act_is_search[a]) {
          out->query_key[q] = s->act_key[a];
          out->query_pos[q] = len - (a - a0 + 1);
          out->query_batch[q] = gi;
          ++q;
          ++cur;
        } else {
          out->qp_src[eq] = v_nq + cur;
          out->qp_dst[eq] = v_np + slot_of(uniq, n_uniq, s->act_key[a]);
          ++eq;
        }
      }
      // product nodes, counts, positions grouped by product
      int64_t e = o.ne;
      if (n_uniq == 0) {
        out->product_key[o.np] = 0;
        out->product_cnt[o.np] = 1;
        out->product_batch[o.np] = gi;
        out->product_pos[e] = 0;
        if (out->last_click_mask) out->last_click_mask[o.np] = 1.0f;
      } else {
        for (int u = 0; u < n_uniq; ++u) {
          int64_t cnt = 0;
          for (int64_t a = a0; a < a1; ++a)
            if (!s->act_is_search[a] && s->act_key[a] == uniq[u]) {
              out->product_pos[e++] = len - (a - a0);
              ++cnt;
            }
          out->product_key[o.np + u] = uniq[u];
          out->product_cnt[o.np + u] = cnt;
          out->product_batch[o.np + u] = gi;
          if (out->last_click_mask) out->last_click_mask[o.np + u] = 0.0f;
        }
      }
      // p -> p transitions
      bool bad = false;
      const int np_pairs = transitions(s, i, chain, pa, pb, pw, &bad);
      for (int j = 0; j < np_pairs; ++j) {
        out->pp_src[o.epp + j] = v_np + pa[(size_t)j];
        out->pp_dst[o.epp + j] = v_np + pb[(size_t)j];
        if (out->pp_weight) out->pp_weight[o.epp + j] = (float)pw[(size_t)j];
      }
      if (out->last_click_mask && n_uniq > 0)
        out->last_click_mask[o.np + (chain.size() >= 2 ? chain.back() : 0)] = 1.0f;
    }
  };
  parallel(fill);
  return 0;
}

extern "C" int sss_featurize_batch(const sss_flat_sessions_t* s, int64_t root_query_key, sss_graph_arrays_t* out,
                                   int n_threads) {
  return featurize_impl(s, 0, root_query_key, out, nullptr, n_threads);
}

extern "C" int sss_featurize_batches(const sss_flat_sessions_t* s, int64_t batch_size, int64_t root_query_key,
                                     sss_graph_arrays_t* out, int64_t* bounds, int n_threads) {
  SSS_REQUIRE(batch_size >= 1 && bounds, "sss_featurize_batches: bad argument");
  return featurize_impl(s, batch_size, root_query_key, out, bounds, n_threads);
}
