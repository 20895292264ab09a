/*
 * hnsw_oracle.c -- CPU restatement of the HNSW branch of HnswIndex::search
 * (/root/reference/crates/cortex-core/src/vector/index.rs:342-373, 416-435).
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see cortex_oracle.c).  It exists to report the
 * reference's approximate path as recall@k against the exact scan, as the north star asks.
 *
 * PARITY UNPINNED.  The graph lives in the third-party crate `instant-distance 0.6.1`
 * (Cargo.lock:2066-2069), which is not vendored under /root/reference and cannot be
 * fetched (no network).  This file restates the PUBLISHED algorithm of that crate's
 * lineage -- Malkov & Yashunin, "Efficient and robust approximate nearest neighbor search
 * using Hierarchical Navigable Small World graphs" (Algorithms 1-5) -- with the parameters
 * the crate's `Builder::default()` is believed to use (M = 32 neighbours per upper layer,
 * 2M at layer 0, ef_construction = 100, ef_search = 100, level multiplier 1/ln(M),
 * heuristic neighbour selection without candidate extension, keeping pruned connections).
 * Those defaults are from memory and UNVERIFIED; the reference's own docs disagree with its
 * code about them (ARCHITECTURE.md:81-84 vs index.rs:430).  The crate seeds its level
 * generator randomly, so the reference itself is not reproducible run to run; no reference
 * test pins recall, parameters or any score for this path.
 *
 * The distance callback at QUERY time is the reference's own (cxo_distance = index.rs:169-179), so
 * the scores of whatever ids come back are reference arithmetic and queries/s is what a scalar
 * per-pair distance costs.  The BUILD is a courtesy: it runs on all host threads (the crate builds with
 * rayon) and uses hoisted norms and a vectorised dot product, otherwise a million rows would not finish
 * inside the benchmark; neighbour lists are kept sorted by distance and a back-link is a sorted insert
 * that drops the farthest entry (as the crate is believed to do), not a second pruning pass.  The graph
 * is therefore "an" instant-distance-style graph, not "the" graph -- which the reference cannot
 * reproduce either.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

float cxo_distance(const float *a, size_t na, const float *b, size_t nb);

typedef struct {
  uint32_t id;
  float d;
} hn_cand;

typedef struct cxo_hnsw {
  const float *vecs; /* n x dim, borrowed */
  size_t n, dim;
  int M, M0, efc, efs;
  int *level;        /* per node */
  hn_cand **nbr;     /* nbr[node]: layer 0 occupies M0 slots, upper layers M each; sorted by distance to the node */
  uint16_t **cnt;    /* cnt[node][layer] */
  unsigned char *lock; /* per-node spin lock (build only) */
  float *rnorm;      /* 1 / |row| for the build-time distance */
  int top_level;
  uint32_t entry;
  uint64_t n_dist;   /* distance evaluations at query time (for reporting) */
  uint32_t *visit;   /* visit stamps of the (single-threaded) query path */
  uint32_t stamp;
} cxo_hnsw;

/* ---- distances ------------------------------------------------------------------------------- */
/* build time: 1 - dot / (|a| |b|) with sixty-four independent partial sums (vectorisable without
 * reassociating any single sum); clones for the vector ISAs the host may have */
__attribute__((target_clones("avx512f", "avx2", "default")))
static float dot16(const float *a, const float *b, size_t n) {
  float acc[64] = {0};
  size_t i = 0;
  for (; i + 64 <= n; i += 64)
    for (int j = 0; j < 64; ++j) acc[j] += a[i + j] * b[i + j];
  for (; i + 16 <= n; i += 16)
    for (int j = 0; j < 16; ++j) acc[j] += a[i + j] * b[i + j];
  float s = 0.0f;
  for (; i < n; ++i) s += a[i] * b[i];
  for (int j = 0; j < 64; ++j) s += acc[j];
  return s;
}

static float dist_build(const cxo_hnsw *h, uint32_t a, uint32_t b) {
  const float d = dot16(h->vecs + (size_t)a * h->dim, h->vecs + (size_t)b * h->dim, h->dim);
  return 1.0f - d * h->rnorm[a] * h->rnorm[b];
}

/* query time: the reference's own arithmetic (index.rs:169-179) */
static float dist_query(cxo_hnsw *h, const float *q, uint32_t b) {
  h->n_dist++;
  return cxo_distance(q, h->dim, h->vecs + (size_t)b * h->dim, h->dim);
}

static int cap_of(const cxo_hnsw *h, int layer) { return layer == 0 ? h->M0 : h->M; }
static hn_cand *nb_of(const cxo_hnsw *h, uint32_t node, int layer) {
  return h->nbr[node] + (layer == 0 ? 0 : h->M0 + (layer - 1) * h->M);
}
static void lock_node(cxo_hnsw *h, uint32_t v) {
  while (__atomic_test_and_set(&h->lock[v], __ATOMIC_ACQUIRE)) {
  }
}
static void unlock_node(cxo_hnsw *h, uint32_t v) { __atomic_clear(&h->lock[v], __ATOMIC_RELEASE); }

/* binary heaps over hn_cand */
typedef struct {
  hn_cand *a;
  int n, cap, max_heap;
} heap;
static void heap_init(heap *p, int cap, int max_heap) {
  p->a = (hn_cand *)malloc((size_t)(cap + 1) * sizeof(hn_cand));
  p->n = 0;
  p->cap = cap;
  p->max_heap = max_heap;
}
static int heap_before(const heap *p, hn_cand x, hn_cand y) { return p->max_heap ? x.d > y.d : x.d < y.d; }
static void heap_push(heap *p, hn_cand c) {
  if (p->n == p->cap) {
    p->cap *= 2;
    p->a = (hn_cand *)realloc(p->a, (size_t)(p->cap + 1) * sizeof(hn_cand));
  }
  int i = p->n++;
  p->a[i] = c;
  while (i > 0) {
    int par = (i - 1) / 2;
    if (!heap_before(p, p->a[i], p->a[par])) break;
    hn_cand t = p->a[i];
    p->a[i] = p->a[par];
    p->a[par] = t;
    i = par;
  }
}
static hn_cand heap_pop(heap *p) {
  hn_cand top = p->a[0];
  p->a[0] = p->a[--p->n];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < p->n && heap_before(p, p->a[l], p->a[m])) m = l;
    if (r < p->n && heap_before(p, p->a[r], p->a[m])) m = r;
    if (m == i) break;
    hn_cand t = p->a[i];
    p->a[i] = p->a[m];
    p->a[m] = t;
    i = m;
  }
  return top;
}

/* Algorithm 2: SEARCH-LAYER.  Returns up to ef nearest in `out` (unsorted), count as result.
 * self = the node being inserted (build) or UINT32_MAX with q = the query vector (search). */
static int search_layer(cxo_hnsw *h, const float *q, uint32_t self, uint32_t ep, float ep_d, int ef, int layer,
                        hn_cand *out, uint32_t *visit, uint32_t *stamp) {
  heap cand, res;
  heap_init(&cand, 64, 0);
  heap_init(&res, ef + 1, 1);
  if (++*stamp == 0) {
    memset(visit, 0, h->n * sizeof(uint32_t));
    *stamp = 1;
  }
  visit[ep] = *stamp;
  hn_cand e = {ep, ep_d};
  heap_push(&cand, e);
  heap_push(&res, e);
  uint32_t local[256];
  while (cand.n) {
    hn_cand c = heap_pop(&cand);
    if (res.n >= ef && c.d > res.a[0].d) break;
    int m;
    if (self != UINT32_MAX) { /* build: other threads edit the lists */
      lock_node(h, c.id);
      const hn_cand *nb = nb_of(h, c.id, layer);
      m = h->cnt[c.id][layer];
      for (int i = 0; i < m; ++i) local[i] = nb[i].id;
      unlock_node(h, c.id);
    } else {
      const hn_cand *nb = nb_of(h, c.id, layer);
      m = h->cnt[c.id][layer];
      for (int i = 0; i < m; ++i) local[i] = nb[i].id;
    }
    for (int i = 0; i < m; ++i) {
      uint32_t v = local[i];
      if (visit[v] == *stamp) continue;
      visit[v] = *stamp;
      float d = self != UINT32_MAX ? dist_build(h, self, v) : dist_query(h, q, v);
      if (res.n < ef || d < res.a[0].d) {
        hn_cand x = {v, d};
        heap_push(&cand, x);
        heap_push(&res, x);
        if (res.n > ef) heap_pop(&res);
      }
    }
  }
  int n = res.n;
  memcpy(out, res.a, (size_t)n * sizeof(hn_cand));
  free(cand.a);
  free(res.a);
  return n;
}

static int cmp_cand(const void *a, const void *b) {
  float x = ((const hn_cand *)a)->d, y = ((const hn_cand *)b)->d;
  return (x > y) - (x < y);
}

/* Algorithm 4: SELECT-NEIGHBORS-HEURISTIC (no candidate extension, keep pruned connections) */
static int select_heuristic(cxo_hnsw *h, hn_cand *c, int n, int M, hn_cand *out) {
  qsort(c, (size_t)n, sizeof(hn_cand), cmp_cand);
  int m = 0, nd = 0;
  hn_cand *disc = (hn_cand *)malloc((size_t)(n > 0 ? n : 1) * sizeof(hn_cand));
  for (int i = 0; i < n && m < M; ++i) {
    int good = 1;
    for (int j = 0; j < m; ++j) {
      if (dist_build(h, c[i].id, out[j].id) < c[i].d) {
        good = 0;
        break;
      }
    }
    if (good) out[m++] = c[i];
    else disc[nd++] = c[i];
  }
  for (int i = 0; i < nd && m < M; ++i) out[m++] = disc[i]; /* keepPrunedConnections */
  free(disc);
  return m;
}

/* keep `list` (cnt entries, capacity cap) sorted by distance: insert x, dropping the farthest when full.
 * The caller holds the owner's lock. */
static int sorted_insert(hn_cand *list, int cnt, int cap, hn_cand x) {
  for (int i = 0; i < cnt; ++i)
    if (list[i].id == x.id) return cnt;
  int lo = 0, hi = cnt;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (list[mid].d <= x.d) lo = mid + 1;
    else hi = mid;
  }
  if (lo >= cap) return cnt;
  int last = cnt < cap ? cnt : cap - 1;
  for (int i = last; i > lo; --i) list[i] = list[i - 1];
  list[lo] = x;
  return cnt < cap ? cnt + 1 : cap;
}

static void insert_node(cxo_hnsw *h, uint32_t i, uint32_t *visit, uint32_t *stamp, hn_cand *W, hn_cand *sel) {
  const int lvl = h->level[i];
  uint32_t ep;
  int top;
#pragma omp critical(cxo_hnsw_entry)
  {
    ep = h->entry;
    top = h->top_level;
  }
  float ep_d = dist_build(h, i, ep);
  uint32_t local[256];
  for (int l = top; l > lvl; --l) { /* greedy descent, ef = 1 */
    int changed = 1;
    while (changed) {
      changed = 0;
      lock_node(h, ep);
      const hn_cand *nb = nb_of(h, ep, l);
      int m = h->cnt[ep][l];
      for (int j = 0; j < m; ++j) local[j] = nb[j].id;
      unlock_node(h, ep);
      for (int j = 0; j < m; ++j) {
        float d = dist_build(h, i, local[j]);
        if (d < ep_d) {
          ep_d = d;
          ep = local[j];
          changed = 1;
        }
      }
    }
  }
  for (int l = lvl < top ? lvl : top; l >= 0; --l) {
    int nw = search_layer(h, NULL, i, ep, ep_d, h->efc, l, W, visit, stamp);
    int m = select_heuristic(h, W, nw, h->M, sel); /* W is sorted ascending on return */
    lock_node(h, i);
    {
      hn_cand *mine = nb_of(h, i, l);
      int c = h->cnt[i][l];
      for (int j = 0; j < m; ++j) c = sorted_insert(mine, c, cap_of(h, l), sel[j]);
      h->cnt[i][l] = (uint16_t)c;
    }
    unlock_node(h, i);
    for (int j = 0; j < m; ++j) { /* back-links: sorted insert, the farthest entry gives way */
      const uint32_t v = sel[j].id;
      hn_cand x = {i, sel[j].d};
      lock_node(h, v);
      h->cnt[v][l] = (uint16_t)sorted_insert(nb_of(h, v, l), h->cnt[v][l], cap_of(h, l), x);
      unlock_node(h, v);
    }
    ep = W[0].id; /* closest found becomes the entry point of the next layer */
    ep_d = W[0].d;
  }
  if (lvl > top) {
#pragma omp critical(cxo_hnsw_entry)
    {
      if (lvl > h->top_level) {
        h->top_level = lvl;
        h->entry = i;
      }
    }
  }
}

cxo_hnsw *cxo_hnsw_build_mt(const float *vecs, size_t n, size_t dim, int M, int efc, int efs, uint64_t seed,
                            int n_threads) {
  cxo_hnsw *h = (cxo_hnsw *)calloc(1, sizeof(cxo_hnsw));
  h->vecs = vecs;
  h->n = n;
  h->dim = dim;
  h->M = M > 1 ? M : 32;
  if (h->M > 120) h->M = 120;
  h->M0 = 2 * h->M;
  h->efc = efc > 0 ? efc : 100;
  h->efs = efs > 0 ? efs : 100;
  uint64_t rng = seed ? seed : 0x9E3779B97F4A7C15ull;
  h->level = (int *)calloc(n ? n : 1, sizeof(int));
  h->nbr = (hn_cand **)calloc(n ? n : 1, sizeof(hn_cand *));
  h->cnt = (uint16_t **)calloc(n ? n : 1, sizeof(uint16_t *));
  h->lock = (unsigned char *)calloc(n ? n : 1, 1);
  h->rnorm = (float *)calloc(n ? n : 1, sizeof(float));
  h->visit = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  h->top_level = -1;
  const double ml = 1.0 / log((double)h->M);
  for (size_t i = 0; i < n; ++i) { /* levels first, in order: the same seed gives the same level assignment */
    rng ^= rng >> 12;
    rng ^= rng << 25;
    rng ^= rng >> 27;
    double u = (double)((rng * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
    if (u < 1e-300) u = 1e-300;
    int lvl = (int)floor(-log(u) * ml);
    if (lvl > 30) lvl = 30;
    h->level[i] = lvl;
    h->nbr[i] = (hn_cand *)malloc((size_t)(h->M0 + lvl * h->M) * sizeof(hn_cand));
    h->cnt[i] = (uint16_t *)calloc((size_t)lvl + 1, sizeof(uint16_t));
  }
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#else
  (void)n_threads;
#endif
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) {
    const float *v = vecs + (size_t)i * dim;
    float ss = dot16(v, v, dim);
    h->rnorm[i] = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
  }
  if (!n) return h;
  h->top_level = h->level[0];
  h->entry = 0;
#pragma omp parallel
  {
    uint32_t *visit = (uint32_t *)calloc(n, sizeof(uint32_t));
    uint32_t stamp = 0;
    hn_cand *W = (hn_cand *)malloc((size_t)(h->efc + 2) * sizeof(hn_cand));
    hn_cand *sel = (hn_cand *)malloc((size_t)(h->M0 + 2) * sizeof(hn_cand));
#pragma omp for schedule(dynamic, 32)
    for (long i = 1; i < (long)n; ++i) insert_node(h, (uint32_t)i, visit, &stamp, W, sel);
    free(visit);
    free(W);
    free(sel);
  }
  return h;
}

cxo_hnsw *cxo_hnsw_build(const float *vecs, size_t n, size_t dim, int M, int efc, int efs, uint64_t seed) {
  return cxo_hnsw_build_mt(vecs, n, dim, M, efc, efs, seed, 1);
}

/* HnswMap::search with Search::default(): ascending distance, at most ef_search results.
 * Returns the number written (<= max_out).  Single-threaded per handle (visit stamps). */
size_t cxo_hnsw_search(cxo_hnsw *h, const float *q, size_t max_out, uint32_t *out_ids, float *out_dist) {
  if (!h->n) return 0;
  uint32_t ep = h->entry;
  float ep_d = dist_query(h, q, ep);
  for (int l = h->top_level; l > 0; --l) {
    int changed = 1;
    while (changed) {
      changed = 0;
      const hn_cand *nb = nb_of(h, ep, l);
      int m = h->cnt[ep][l];
      for (int j = 0; j < m; ++j) {
        float d = dist_query(h, q, nb[j].id);
        if (d < ep_d) {
          ep_d = d;
          ep = nb[j].id;
          changed = 1;
        }
      }
    }
  }
  hn_cand *W = (hn_cand *)malloc((size_t)(h->efs + 2) * sizeof(hn_cand));
  int nw = search_layer(h, q, UINT32_MAX, ep, ep_d, h->efs, 0, W, h->visit, &h->stamp);
  qsort(W, (size_t)nw, sizeof(hn_cand), cmp_cand);
  size_t n = (size_t)nw < max_out ? (size_t)nw : max_out;
  for (size_t i = 0; i < n; ++i) {
    out_ids[i] = W[i].id;
    out_dist[i] = W[i].d;
  }
  free(W);
  return n;
}

uint64_t cxo_hnsw_distance_evals(const cxo_hnsw *h) { return h->n_dist; }

void cxo_hnsw_free(cxo_hnsw *h) {
  if (!h) return;
  for (size_t i = 0; i < h->n; ++i) {
    free(h->nbr[i]);
    free(h->cnt[i]);
  }
  free(h->nbr);
  free(h->cnt);
  free(h->level);
  free(h->lock);
  free(h->rnorm);
  free(h->visit);
  free(h);
}
