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
 * The distance callback is the reference's own (cxo_distance = index.rs:169-179), so the
 * scores of whatever ids come back are reference arithmetic.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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
  uint32_t **nbr;    /* nbr[node][layer * stride .. ] flattened per node */
  uint16_t **cnt;    /* cnt[node][layer] */
  int top_level;
  uint32_t entry;
  uint64_t rng;
  uint64_t n_dist;   /* distance evaluations (for reporting) */
  uint32_t *visit;   /* visit stamps */
  uint32_t stamp;
} cxo_hnsw;

static double rnd(cxo_hnsw *h) { /* xorshift64* */
  h->rng ^= h->rng >> 12;
  h->rng ^= h->rng << 25;
  h->rng ^= h->rng >> 27;
  return (double)((h->rng * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
}

static float dist(cxo_hnsw *h, const float *q, uint32_t b) {
  h->n_dist++;
  return cxo_distance(q, h->dim, h->vecs + (size_t)b * h->dim, h->dim);
}

static int cap_of(const cxo_hnsw *h, int layer) { return layer == 0 ? h->M0 : h->M; }
static uint32_t *nb_of(const cxo_hnsw *h, uint32_t node, int layer) {
  /* layer 0 occupies M0 slots, upper layers M each */
  return h->nbr[node] + (layer == 0 ? 0 : h->M0 + (layer - 1) * h->M);
}

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

/* Algorithm 2: SEARCH-LAYER.  Returns up to ef nearest in `out` (unsorted), count as result. */
static int search_layer(cxo_hnsw *h, const float *q, uint32_t ep, float ep_d, int ef, int layer, hn_cand *out) {
  heap cand, res;
  heap_init(&cand, 64, 0);
  heap_init(&res, ef + 1, 1);
  if (++h->stamp == 0) {
    memset(h->visit, 0, h->n * sizeof(uint32_t));
    h->stamp = 1;
  }
  h->visit[ep] = h->stamp;
  hn_cand e = {ep, ep_d};
  heap_push(&cand, e);
  heap_push(&res, e);
  while (cand.n) {
    hn_cand c = heap_pop(&cand);
    if (res.n >= ef && c.d > res.a[0].d) break;
    uint32_t *nb = nb_of(h, c.id, layer);
    int m = h->cnt[c.id][layer];
    for (int i = 0; i < m; ++i) {
      uint32_t v = nb[i];
      if (h->visit[v] == h->stamp) continue;
      h->visit[v] = h->stamp;
      float d = dist(h, q, v);
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
static int select_heuristic(cxo_hnsw *h, hn_cand *c, int n, int M, uint32_t *out) {
  qsort(c, (size_t)n, sizeof(hn_cand), cmp_cand);
  int m = 0, nd = 0;
  hn_cand *disc = (hn_cand *)malloc((size_t)(n > 0 ? n : 1) * sizeof(hn_cand));
  for (int i = 0; i < n && m < M; ++i) {
    int good = 1;
    for (int j = 0; j < m; ++j) {
      float dj = dist(h, h->vecs + (size_t)c[i].id * h->dim, out[j]);
      if (dj < c[i].d) {
        good = 0;
        break;
      }
    }
    if (good) out[m++] = c[i].id;
    else disc[nd++] = c[i];
  }
  for (int i = 0; i < nd && m < M; ++i) out[m++] = disc[i].id; /* keepPrunedConnections */
  free(disc);
  return m;
}

cxo_hnsw *cxo_hnsw_build(const float *vecs, size_t n, size_t dim, int M, int efc, int efs, uint64_t seed) {
  cxo_hnsw *h = (cxo_hnsw *)calloc(1, sizeof(cxo_hnsw));
  h->vecs = vecs;
  h->n = n;
  h->dim = dim;
  h->M = M > 1 ? M : 32;
  h->M0 = 2 * h->M;
  h->efc = efc > 0 ? efc : 100;
  h->efs = efs > 0 ? efs : 100;
  h->rng = seed ? seed : 0x9E3779B97F4A7C15ull;
  h->level = (int *)calloc(n ? n : 1, sizeof(int));
  h->nbr = (uint32_t **)calloc(n ? n : 1, sizeof(uint32_t *));
  h->cnt = (uint16_t **)calloc(n ? n : 1, sizeof(uint16_t *));
  h->visit = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  h->top_level = -1;
  const double ml = 1.0 / log((double)h->M);
  hn_cand *W = (hn_cand *)malloc((size_t)(h->efc + 2) * sizeof(hn_cand));
  uint32_t *sel = (uint32_t *)malloc((size_t)h->M0 * sizeof(uint32_t));
  hn_cand *tmp = (hn_cand *)malloc((size_t)(h->M0 + 2) * sizeof(hn_cand));
  for (size_t i = 0; i < n; ++i) {
    double u = rnd(h);
    if (u < 1e-300) u = 1e-300;
    int lvl = (int)floor(-log(u) * ml);
    if (lvl > 30) lvl = 30;
    h->level[i] = lvl;
    h->nbr[i] = (uint32_t *)malloc((size_t)(h->M0 + lvl * h->M) * sizeof(uint32_t));
    h->cnt[i] = (uint16_t *)calloc((size_t)lvl + 1, sizeof(uint16_t));
    const float *q = vecs + i * dim;
    if (h->top_level < 0) {
      h->top_level = lvl;
      h->entry = (uint32_t)i;
      continue;
    }
    uint32_t ep = h->entry;
    float ep_d = dist(h, q, ep);
    for (int l = h->top_level; l > lvl; --l) { /* greedy descent, ef = 1 */
      int changed = 1;
      while (changed) {
        changed = 0;
        uint32_t *nb = nb_of(h, ep, l);
        int m = h->cnt[ep][l];
        for (int j = 0; j < m; ++j) {
          float d = dist(h, q, nb[j]);
          if (d < ep_d) {
            ep_d = d;
            ep = nb[j];
            changed = 1;
          }
        }
      }
    }
    for (int l = lvl < h->top_level ? lvl : h->top_level; l >= 0; --l) {
      int nw = search_layer(h, q, ep, ep_d, h->efc, l, W);
      int m = select_heuristic(h, W, nw, h->M, sel); /* W is sorted ascending on return */
      uint32_t *mine = nb_of(h, (uint32_t)i, l);
      memcpy(mine, sel, (size_t)m * sizeof(uint32_t));
      h->cnt[i][l] = (uint16_t)m;
      for (int j = 0; j < m; ++j) { /* bidirectional links, shrink with the same heuristic */
        uint32_t v = sel[j];
        uint32_t *vn = nb_of(h, v, l);
        int vc = h->cnt[v][l], cap = cap_of(h, l);
        if (vc < cap) {
          vn[vc] = (uint32_t)i;
          h->cnt[v][l] = (uint16_t)(vc + 1);
        } else {
          const float *vq = vecs + (size_t)v * dim;
          for (int t = 0; t < vc; ++t) {
            tmp[t].id = vn[t];
            tmp[t].d = dist(h, vq, vn[t]);
          }
          tmp[vc].id = (uint32_t)i;
          tmp[vc].d = dist(h, vq, (uint32_t)i);
          h->cnt[v][l] = (uint16_t)select_heuristic(h, tmp, vc + 1, cap, vn);
        }
      }
      ep = W[0].id; /* closest found becomes the entry point of the next layer */
      ep_d = W[0].d;
    }
    if (lvl > h->top_level) {
      h->top_level = lvl;
      h->entry = (uint32_t)i;
    }
  }
  free(W);
  free(sel);
  free(tmp);
  return h;
}

/* HnswMap::search with Search::default(): ascending distance, at most ef_search results.
 * Returns the number written (<= max_out). */
size_t cxo_hnsw_search(cxo_hnsw *h, const float *q, size_t max_out, uint32_t *out_ids, float *out_dist) {
  if (!h->n) return 0;
  uint32_t ep = h->entry;
  float ep_d = dist(h, q, ep);
  for (int l = h->top_level; l > 0; --l) {
    int changed = 1;
    while (changed) {
      changed = 0;
      uint32_t *nb = nb_of(h, ep, l);
      int m = h->cnt[ep][l];
      for (int j = 0; j < m; ++j) {
        float d = dist(h, q, nb[j]);
        if (d < ep_d) {
          ep_d = d;
          ep = nb[j];
          changed = 1;
        }
      }
    }
  }
  hn_cand *W = (hn_cand *)malloc((size_t)(h->efs + 2) * sizeof(hn_cand));
  int nw = search_layer(h, q, ep, ep_d, h->efs, 0, W);
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
  free(h->visit);
  free(h);
}
