/*
 * search_oracle.c — CPU restatement ("oracle O2") of the reference's brute-force retrieval path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may call this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
 *
 * What it restates (file:line into the reference tree):
 *   normalize()                       util_amazon_filtered.py:28-31, fine_tune_ours.py:38-40
 *   build_index + IndexFlatIP/L2      test_amazon_filterd.py:207-223 (faiss, not vendored; unpinned in
 *                                     dependency.txt:3) — exact fp32 inner product / squared L2,
 *                                     results best-first
 *   IndexBinaryFlat                   fine_tune_ours.py:839-843,871-876 — Hamming distance, ascending
 *   get_prediction_by_knn             test_amazon_filterd.py:59-78 — per-item sum of neighbour weights
 *   segment max / sum    defined by this repo, not by the reference
 *
 * Parity status: the reference ships no golden vectors for this path and faiss is not
 * installable here, so this oracle is pinned against (a) the reference's own normalize() imported from
 * /root/reference through a module shim (tests/golden/gen_golden.py) and (b) a numpy/float64 brute force;
 * the faiss boundary itself is "parity unpinned".
 *
 * Arithmetic is made well defined where faiss/BLAS leave it open:
 *   score = fmaf(q[j], x[j], acc) for j = 0..d-1 (k-ascending, one rounding per step, acc starts at 0)
 *   L2    = fmaf(q[j]-x[j], q[j]-x[j], acc)
 *   order = (score desc | distance asc, then id asc)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define O_NORM_NONE 0
#define O_NORM_UTIL 1
#define O_NORM_FT 2
#define O_NORM_TORCH 3

/* util_amazon_filtered.py:28-31 / fine_tune_ours.py:38-40 / F.normalize, with a fixed summation order */
void o_normalize(const float* in, float* out, int64_t n, int d, int mode) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) { /* rows are independent; the order INSIDE a row is what is fixed */
    const float* x = in + i * (int64_t)d;
    float* y = out + i * (int64_t)d;
    if (mode == O_NORM_NONE) {
      if (y != x) memcpy(y, x, sizeof(float) * (size_t)d);
      continue;
    }
    float ss = 0.0f;
    for (int j = 0; j < d; ++j) ss = fmaf(x[j], x[j], ss);
    float den;
    if (mode == O_NORM_UTIL) {
      den = sqrtf(ss < 1e-6f ? 1e-6f : ss);
    } else if (mode == O_NORM_FT) {
      den = sqrtf(ss) + 1e-4f;
    } else {
      float nn = sqrtf(ss);
      den = nn < 1e-12f ? 1e-12f : nn;
    }
    for (int j = 0; j < d; ++j) y[j] = x[j] / den;
  }
}

static inline float dot_fixed(const float* q, const float* x, int d) {
  float acc = 0.0f;
  for (int j = 0; j < d; ++j) acc = fmaf(q[j], x[j], acc);
  return acc;
}
static inline float l2_fixed(const float* q, const float* x, int d) {
  float acc = 0.0f;
  for (int j = 0; j < d; ++j) {
    float t = q[j] - x[j];
    acc = fmaf(t, t, acc);
  }
  return acc;
}

/* "a ranks before b": larger score first, then smaller id */
typedef struct {
  float s;
  int64_t id;
} ent_t;
static inline int before(ent_t a, ent_t b) { return a.s > b.s || (a.s == b.s && a.id < b.id); }

/* heap with the worst retained element at the root */
static void heap_push(ent_t* h, int* n, int k, ent_t e) {
  if (*n < k) {
    int i = (*n)++;
    h[i] = e;
    while (i > 0) {
      int p = (i - 1) / 2;
      if (before(h[p], h[i])) { /* parent better than child -> child must go up (root = worst) */
        ent_t t = h[p];
        h[p] = h[i];
        h[i] = t;
        i = p;
      } else
        break;
    }
  } else if (before(e, h[0])) {
    h[0] = e;
    int i = 0;
    for (;;) {
      int l = 2 * i + 1, r = l + 1, w = i;
      if (l < k && before(h[w], h[l])) w = l;
      if (r < k && before(h[w], h[r])) w = r;
      if (w == i) break;
      ent_t t = h[w];
      h[w] = h[i];
      h[i] = t;
      i = w;
    }
  }
}
static int cmp_ent(const void* a, const void* b) {
  ent_t x = *(const ent_t*)a, y = *(const ent_t*)b;
  if (before(x, y)) return -1;
  if (before(y, x)) return 1;
  return 0;
}

/*
 * Flat search.  metric 0 = IP, 1 = L2.  reduce 0 = none, 1 = max, 2 = sum (seg_off[n_seg+1] required for
 * 1 and 2; ids are then segment indices).  D [nq,k], I [nq,k]; tail padded with (-inf | +inf, -1).
 * id_offset is added to returned ids.
 */
int o_search_flat(const float* db, int64_t n, int d, const float* q, int64_t nq, int k, int metric,
                  const int64_t* seg_off, int64_t n_seg, int reduce, int64_t id_offset, float* D, int64_t* I) {
  if (reduce != 0 && !seg_off) return 1;
  float* sumrows = NULL;
  if (reduce == 2) {
    /* sum over a segment is linear: score = <q, sum of rows>, rows added in row order */
    sumrows = (float*)calloc((size_t)n_seg * (size_t)d, sizeof(float));
    if (!sumrows) return 2;
    for (int64_t s = 0; s < n_seg; ++s)
      for (int64_t r = seg_off[s]; r < seg_off[s + 1]; ++r)
        for (int j = 0; j < d; ++j) sumrows[s * d + j] += db[r * d + j];
  }
  int64_t n_units = reduce == 0 ? n : n_seg;
#pragma omp parallel
  {
    ent_t* h = (ent_t*)malloc(sizeof(ent_t) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; ++qi) {
      const float* qv = q + qi * (int64_t)d;
      int hn = 0;
      for (int64_t u = 0; u < n_units; ++u) {
        float s;
        if (reduce == 0) {
          s = metric == 0 ? dot_fixed(qv, db + u * d, d) : -l2_fixed(qv, db + u * d, d);
        } else if (reduce == 2) {
          s = metric == 0 ? dot_fixed(qv, sumrows + u * d, d) : -l2_fixed(qv, sumrows + u * d, d);
        } else {
          if (seg_off[u + 1] == seg_off[u]) continue; /* empty segment has no score */
          s = -INFINITY;
          for (int64_t r = seg_off[u]; r < seg_off[u + 1]; ++r) {
            float t = metric == 0 ? dot_fixed(qv, db + r * d, d) : -l2_fixed(qv, db + r * d, d);
            if (t > s) s = t;
          }
        }
        ent_t e = {s, u};
        heap_push(h, &hn, k, e);
      }
      qsort(h, (size_t)hn, sizeof(ent_t), cmp_ent);
      for (int j = 0; j < k; ++j) {
        if (j < hn) {
          D[qi * k + j] = metric == 0 ? h[j].s : -h[j].s;
          I[qi * k + j] = h[j].id + id_offset;
        } else {
          D[qi * k + j] = metric == 0 ? -INFINITY : INFINITY;
          I[qi * k + j] = -1;
        }
      }
    }
    free(h);
  }
  free(sumrows);
  return 0;
}

/* Hamming top-k over packed codes (fine_tune_ours.py:839-843,871-876): distance ascending, id ascending */
int o_search_hamming(const uint8_t* db, int64_t n, int nbytes, const uint8_t* q, int64_t nq, int k,
                     int64_t id_offset, int32_t* D, int64_t* I) {
#pragma omp parallel
  {
    ent_t* h = (ent_t*)malloc(sizeof(ent_t) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; ++qi) {
      const uint8_t* qv = q + qi * (int64_t)nbytes;
      int hn = 0;
      for (int64_t u = 0; u < n; ++u) {
        const uint8_t* x = db + u * (int64_t)nbytes;
        int dist = 0;
        for (int j = 0; j < nbytes; ++j) dist += __builtin_popcount((unsigned)(qv[j] ^ x[j]));
        ent_t e = {-(float)dist, u};
        heap_push(h, &hn, k, e);
      }
      qsort(h, (size_t)hn, sizeof(ent_t), cmp_ent);
      for (int j = 0; j < k; ++j) {
        if (j < hn) {
          D[qi * k + j] = (int32_t)(-h[j].s);
          I[qi * k + j] = h[j].id + id_offset;
        } else {
          D[qi * k + j] = INT32_MAX;
          I[qi * k + j] = -1;
        }
      }
    }
    free(h);
  }
  return 0;
}

/* sign-binarise and pack MSB first, zero padded (fine_tune_ours.py:839-840 on model/model.py:137 output) */
void o_pack_sign_bits(const float* x, uint8_t* codes, int64_t n, int nbits) {
  int nbytes = (nbits + 7) / 8;
  for (int64_t i = 0; i < n; ++i)
    for (int b = 0; b < nbytes; ++b) {
      unsigned v = 0;
      for (int j = 0; j < 8; ++j) {
        int col = b * 8 + j;
        v = (v << 1) | (unsigned)(col < nbits && x[i * (int64_t)nbits + col] > 0.0f);
      }
      codes[i * (int64_t)nbytes + b] = (uint8_t)v;
    }
}

/*
 * get_prediction_by_knn (test_amazon_filterd.py:59-78): item weight = sum over neighbour sessions (in
 * arrival order, accumulated in float64: `np.ones_like(int64 ids) * np.float32` is a float64 array, :71) of that
 * neighbour's similarity; top-K by weight descending with Python's STABLE sort (:76), i.e. equal weights keep
 * the order in which the items first arrived in the defaultdict (:72-74).  Pinned by tests/golden/vote_golden.npz,
 * which the reference's own function produced.
 */
typedef struct {
  int64_t item;
  double w;
  int64_t first;
} vote_t;
static int cmp_item(const void* a, const void* b) {
  int64_t x = ((const vote_t*)a)->item, y = ((const vote_t*)b)->item;
  if (x != y) return x < y ? -1 : 1;
  int64_t fx = ((const vote_t*)a)->first, fy = ((const vote_t*)b)->first;
  return fx < fy ? -1 : (fx > fy ? 1 : 0);
}
static int cmp_vote(const void* a, const void* b) {
  const vote_t *x = (const vote_t*)a, *y = (const vote_t*)b;
  if (x->w != y->w) return x->w > y->w ? -1 : 1;
  return x->first < y->first ? -1 : (x->first > y->first ? 1 : 0);
}
int o_item_vote(const float* D, const int64_t* I, int64_t nq, int s, const int64_t* item_off, const int64_t* items,
                int K, int64_t* out_items, float* out_w) {
  for (int64_t qi = 0; qi < nq; ++qi) {
    int64_t total = 0;
    for (int j = 0; j < s; ++j) {
      int64_t sess = I[qi * s + j];
      if (sess >= 0) total += item_off[sess + 1] - item_off[sess];
    }
    vote_t* v = (vote_t*)malloc(sizeof(vote_t) * (size_t)(total > 0 ? total : 1));
    int64_t m = 0;
    for (int j = 0; j < s; ++j) {
      int64_t sess = I[qi * s + j];
      if (sess < 0) continue;
      for (int64_t t = item_off[sess]; t < item_off[sess + 1]; ++t) {
        v[m].item = items[t];
        v[m].w = D[qi * s + j];
        v[m].first = m;
        ++m;
      }
    }
    qsort(v, (size_t)m, sizeof(vote_t), cmp_item); /* stable by construction: (item, arrival order) */
    int64_t u = 0;
    for (int64_t a = 0; a < m;) {
      int64_t b = a;
      double w = 0.0;
      while (b < m && v[b].item == v[a].item) w += v[b++].w; /* arrival order, float64 */
      v[u].item = v[a].item;
      v[u].first = v[a].first;                               /* (a is the run head: the smallest arrival index) */
      v[u].w = w;
      ++u;
      a = b;
    }
    qsort(v, (size_t)u, sizeof(vote_t), cmp_vote);
    for (int j = 0; j < K; ++j) {
      out_items[qi * K + j] = j < u ? v[j].item : -1;
      out_w[qi * K + j] = j < u ? (float)v[j].w : 0.0f;
    }
    free(v);
  }
  return 0;
}

/* k-way merge restatement for the sharded path */
int o_topk_merge(const float* cD, const int64_t* cI, int n_shards, int64_t nq, int k, int metric, float* D,
                 int64_t* I) {
  ent_t* v = (ent_t*)malloc(sizeof(ent_t) * (size_t)n_shards * (size_t)k);
  for (int64_t qi = 0; qi < nq; ++qi) {
    int m = 0;
    for (int s = 0; s < n_shards; ++s)
      for (int j = 0; j < k; ++j) {
        int64_t id = cI[((int64_t)s * nq + qi) * k + j];
        if (id < 0) continue;
        float sc = cD[((int64_t)s * nq + qi) * k + j];
        v[m].s = metric == 0 ? sc : -sc;
        v[m].id = id;
        ++m;
      }
    qsort(v, (size_t)m, sizeof(ent_t), cmp_ent);
    for (int j = 0; j < k; ++j) {
      if (j < m) {
        D[qi * k + j] = metric == 0 ? v[j].s : -v[j].s;
        I[qi * k + j] = v[j].id;
      } else {
        D[qi * k + j] = metric == 0 ? -INFINITY : INFINITY;
        I[qi * k + j] = -1;
      }
    }
  }
  free(v);
  return 0;
}
