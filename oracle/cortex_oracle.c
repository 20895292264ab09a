/*
 * cortex_oracle.c -- CPU restatement of cortex-core's exact similarity scan.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path.  Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl
 * reference legs of bench.py may load this library, and only as the checker
 * or as the timed CPU baseline.  The product (cortex_b200/) never links,
 * imports or falls back to it.
 *
 * What it restates (all paths relative to /root/reference/):
 *   cxo_distance               crates/cortex-core/src/vector/index.rs:169-179
 *   cxo_distance_to_similarity crates/cortex-core/src/vector/index.rs:253-256
 *   matches_filter             crates/cortex-core/src/vector/index.rs:225-251
 *   cxo_search (brute force)   crates/cortex-core/src/vector/index.rs:259-294, 325-340
 *   cxo_search_threshold       crates/cortex-core/src/vector/index.rs:376-388
 *   cxo_search_batch           crates/cortex-core/src/vector/index.rs:390-410
 *   cxo_insert / remove / len  crates/cortex-core/src/vector/index.rs:298-323, 412-414
 *   cxo_set_metadata           crates/cortex-core/src/vector/index.rs:219-222
 *   cxo_save / cxo_load        crates/cortex-core/src/vector/index.rs:437-472 (bincode 1.3 layout)
 *
 * Parity status: the reference is Rust and no Rust toolchain exists in this
 * image, so the reference itself cannot be executed here.  The restatement is
 * pinned against every assertion the reference's own unit tests make for this
 * path (vector/index.rs:475-729, vector/config.rs:89-135, linker/rules.rs
 * thresholds) in tests/test_oracle_reference_pins.py, and against an
 * independent numpy float32 restatement (tests/golden/).  The reference holds
 * no numeric golden scores for this path, so numeric parity beyond those pins
 * rests on this file following the reference arithmetic operation by
 * operation: three strictly sequential left-to-right fp32 sums of separately
 * rounded products (Rust's Iterator::sum, no FMA contraction, no
 * reassociation), sqrt, one division, two subtractions and a clamp.
 * Build with -ffp-contract=off and without -ffast-math (see Makefile).
 *
 * Two behaviours the reference leaves unspecified are pinned here and in the
 * CUDA path identically, and documented in DESIGN.md:
 *   (1) order among equal scores: the reference iterates a HashMap (random
 *       order) and stable-sorts, so ties come out in arbitrary order.  Here
 *       ties come out in row order (insertion order of live rows).
 *   (2) NaN scores (zero-norm vectors): the reference's comparator maps NaN to
 *       Ordering::Equal, which is not a total order, so their position is
 *       unspecified.  Here NaN-scored rows sort after every non-NaN row, in
 *       row order.  `score >= threshold` is false for NaN in both.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define CXO_OK 0
#define CXO_ERR_VALIDATION 1
#define CXO_ERR_IO 4

/* ------------------------------------------------------------------------ */
/* scalar arithmetic                                                          */

/* vector/index.rs:169-179.  zip() stops at the shorter slice; each norm runs
 * over its own full slice.  volatile-free: -ffp-contract=off keeps mul and add
 * separately rounded. */
float cxo_distance(const float *a, size_t na, const float *b, size_t nb) {
  size_t n = na < nb ? na : nb;
  float dot = 0.0f;
  for (size_t i = 0; i < n; ++i) {
    float p = a[i] * b[i];
    dot = dot + p;
  }
  float sa = 0.0f;
  for (size_t i = 0; i < na; ++i) {
    float p = a[i] * a[i];
    sa = sa + p;
  }
  float sb = 0.0f;
  for (size_t i = 0; i < nb; ++i) {
    float p = b[i] * b[i];
    sb = sb + p;
  }
  float norm_a = sqrtf(sa);
  float norm_b = sqrtf(sb);
  float similarity = dot / (norm_a * norm_b);
  return 1.0f - similarity;
}

/* vector/index.rs:253-256 with Rust's f32::clamp (NaN stays NaN, -0.0 stays). */
float cxo_distance_to_similarity(float distance) {
  float s = 1.0f - distance;
  if (s < 0.0f) s = 0.0f;
  if (s > 1.0f) s = 1.0f;
  return s;
}

/* ------------------------------------------------------------------------ */
/* index state: ordered rows, id hash, tombstones, metadata                   */

typedef struct {
  uint8_t id[16];
  int has_meta;
  uint32_t kind;  /* interned kind string id   */
  uint32_t agent; /* interned agent string id  */
  int live;
} cxo_row;

typedef struct cxo_index {
  size_t dim;
  size_t n_rows; /* rows ever appended (live + dead) */
  size_t n_live;
  size_t cap;
  float *vecs; /* n_rows x dim */
  cxo_row *rows;
  /* open addressing hash: id -> row+1 (0 = empty) */
  uint64_t *slots;
  size_t n_slots;
  /* metadata may be set for an id that has no vector (HashMap semantics) */
  uint8_t *orphan_ids;
  uint32_t *orphan_kind, *orphan_agent;
  size_t n_orphan, cap_orphan;
  /* string tables */
  char **strs;
  size_t n_strs, cap_strs;
  int faithful_copy; /* per-pair row clone like index.rs:270 (baseline timing) */
} cxo_index;

static uint64_t hash_id(const uint8_t *id) {
  uint64_t a, b;
  memcpy(&a, id, 8);
  memcpy(&b, id + 8, 8);
  uint64_t h = a * 0x9E3779B97F4A7C15ull ^ (b + 0xC2B2AE3D27D4EB4Full);
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

static void rehash(cxo_index *ix, size_t n_slots) {
  free(ix->slots);
  ix->slots = (uint64_t *)calloc(n_slots, sizeof(uint64_t));
  ix->n_slots = n_slots;
  for (size_t r = 0; r < ix->n_rows; ++r) {
    if (!ix->rows[r].live) continue;
    size_t s = hash_id(ix->rows[r].id) & (n_slots - 1);
    while (ix->slots[s]) s = (s + 1) & (n_slots - 1);
    ix->slots[s] = r + 1;
  }
}

static long find_row(const cxo_index *ix, const uint8_t *id) {
  if (!ix->n_slots) return -1;
  size_t s = hash_id(id) & (ix->n_slots - 1);
  while (ix->slots[s]) {
    size_t r = ix->slots[s] - 1;
    if (ix->rows[r].live && memcmp(ix->rows[r].id, id, 16) == 0) return (long)r;
    s = (s + 1) & (ix->n_slots - 1);
  }
  return -1;
}

cxo_index *cxo_create(size_t dim) {
  cxo_index *ix = (cxo_index *)calloc(1, sizeof(cxo_index));
  ix->dim = dim;
  ix->faithful_copy = 1;
  return ix;
}

void cxo_destroy(cxo_index *ix) {
  if (!ix) return;
  free(ix->vecs);
  free(ix->rows);
  free(ix->slots);
  free(ix->orphan_ids);
  free(ix->orphan_kind);
  free(ix->orphan_agent);
  for (size_t i = 0; i < ix->n_strs; ++i) free(ix->strs[i]);
  free(ix->strs);
  free(ix);
}

void cxo_set_faithful_copy(cxo_index *ix, int on) { ix->faithful_copy = on; }

uint32_t cxo_intern(cxo_index *ix, const char *s) {
  for (size_t i = 0; i < ix->n_strs; ++i)
    if (strcmp(ix->strs[i], s) == 0) return (uint32_t)i;
  if (ix->n_strs == ix->cap_strs) {
    ix->cap_strs = ix->cap_strs ? ix->cap_strs * 2 : 16;
    ix->strs = (char **)realloc(ix->strs, ix->cap_strs * sizeof(char *));
  }
  ix->strs[ix->n_strs] = strdup(s);
  return (uint32_t)ix->n_strs++;
}

static long find_orphan(const cxo_index *ix, const uint8_t *id) {
  for (size_t i = 0; i < ix->n_orphan; ++i)
    if (memcmp(ix->orphan_ids + 16 * i, id, 16) == 0) return (long)i;
  return -1;
}

/* vector/index.rs:298-314: dimension check, then HashMap::insert (same id
 * overwrites the stored vector; metadata map untouched). */
int cxo_insert(cxo_index *ix, const uint8_t *id, const float *v, size_t len) {
  if (len != ix->dim) return CXO_ERR_VALIDATION;
  long r = find_row(ix, id);
  if (r >= 0) {
    memcpy(ix->vecs + (size_t)r * ix->dim, v, ix->dim * sizeof(float));
    return CXO_OK;
  }
  if (ix->n_rows == ix->cap) {
    ix->cap = ix->cap ? ix->cap * 2 : 64;
    ix->vecs = (float *)realloc(ix->vecs, ix->cap * (ix->dim ? ix->dim : 1) * sizeof(float));
    ix->rows = (cxo_row *)realloc(ix->rows, ix->cap * sizeof(cxo_row));
  }
  size_t row = ix->n_rows++;
  memcpy(ix->vecs + row * ix->dim, v, ix->dim * sizeof(float));
  cxo_row *R = &ix->rows[row];
  memcpy(R->id, id, 16);
  R->live = 1;
  R->has_meta = 0;
  R->kind = R->agent = 0;
  long o = find_orphan(ix, id); /* metadata set before the vector existed */
  if (o >= 0) {
    R->has_meta = 1;
    R->kind = ix->orphan_kind[o];
    R->agent = ix->orphan_agent[o];
    size_t last = --ix->n_orphan;
    memcpy(ix->orphan_ids + 16 * o, ix->orphan_ids + 16 * last, 16);
    ix->orphan_kind[o] = ix->orphan_kind[last];
    ix->orphan_agent[o] = ix->orphan_agent[last];
  }
  ix->n_live++;
  if ((ix->n_rows + 1) * 2 > ix->n_slots) rehash(ix, ix->n_slots ? ix->n_slots * 2 : 128);
  else {
    size_t s = hash_id(id) & (ix->n_slots - 1);
    while (ix->slots[s]) s = (s + 1) & (ix->n_slots - 1);
    ix->slots[s] = row + 1;
  }
  return CXO_OK;
}

/* the startup loop serve.rs:111-117 / api.rs:55-69: one insert per row, in order */
int cxo_insert_batch(cxo_index *ix, const uint8_t *ids, const float *rows, size_t n, size_t len) {
  for (size_t i = 0; i < n; ++i) {
    int rc = cxo_insert(ix, ids + 16 * i, rows + i * len, len);
    if (rc != CXO_OK) return rc;
  }
  return CXO_OK;
}

/* vector/index.rs:316-323: drop vector and metadata; never an error. */
int cxo_remove(cxo_index *ix, const uint8_t *id) {
  long r = find_row(ix, id);
  if (r >= 0) {
    ix->rows[r].live = 0;
    ix->n_live--;
    rehash(ix, ix->n_slots); /* tombstone-free table; removal is rare */
  }
  long o = find_orphan(ix, id);
  if (o >= 0) {
    size_t last = --ix->n_orphan;
    memcpy(ix->orphan_ids + 16 * o, ix->orphan_ids + 16 * last, 16);
    ix->orphan_kind[o] = ix->orphan_kind[last];
    ix->orphan_agent[o] = ix->orphan_agent[last];
  }
  return CXO_OK;
}

/* vector/index.rs:219-222 */
void cxo_set_metadata(cxo_index *ix, const uint8_t *id, const char *kind, const char *agent) {
  uint32_t k = cxo_intern(ix, kind), a = cxo_intern(ix, agent);
  long r = find_row(ix, id);
  if (r >= 0) {
    ix->rows[r].has_meta = 1;
    ix->rows[r].kind = k;
    ix->rows[r].agent = a;
    return;
  }
  long o = find_orphan(ix, id);
  if (o < 0) {
    if (ix->n_orphan == ix->cap_orphan) {
      ix->cap_orphan = ix->cap_orphan ? ix->cap_orphan * 2 : 16;
      ix->orphan_ids = (uint8_t *)realloc(ix->orphan_ids, ix->cap_orphan * 16);
      ix->orphan_kind = (uint32_t *)realloc(ix->orphan_kind, ix->cap_orphan * 4);
      ix->orphan_agent = (uint32_t *)realloc(ix->orphan_agent, ix->cap_orphan * 4);
    }
    o = (long)ix->n_orphan++;
    memcpy(ix->orphan_ids + 16 * o, id, 16);
  }
  ix->orphan_kind[o] = k;
  ix->orphan_agent[o] = a;
}

size_t cxo_len(const cxo_index *ix) { return ix->n_live; }
size_t cxo_dim(const cxo_index *ix) { return ix->dim; }

/* ------------------------------------------------------------------------ */
/* filter                                                                     */

typedef struct {
  int has_kinds;
  const char *const *kinds;
  size_t n_kinds;
  int has_exclude;
  const uint8_t *exclude; /* n_exclude x 16 */
  size_t n_exclude;
  int has_agent;
  const char *agent;
} cxo_filter;

/* vector/index.rs:225-251 */
static int matches_filter(const cxo_index *ix, const cxo_row *R, const cxo_filter *f) {
  if (f->has_exclude)
    for (size_t i = 0; i < f->n_exclude; ++i)
      if (memcmp(f->exclude + 16 * i, R->id, 16) == 0) return 0;
  if (R->has_meta) {
    if (f->has_kinds) {
      int ok = 0;
      for (size_t i = 0; i < f->n_kinds; ++i)
        if (strcmp(f->kinds[i], ix->strs[R->kind]) == 0) ok = 1;
      if (!ok) return 0;
    }
    if (f->has_agent)
      if (strcmp(f->agent, ix->strs[R->agent]) != 0) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------ */
/* brute-force scan                                                           */

typedef struct {
  uint32_t row;
  float score, distance;
} cxo_hit;

/* "b.score.partial_cmp(&a.score)" descending; NaN pinned last (header note 2). */
static int hit_before(const cxo_hit *a, const cxo_hit *b) {
  int an = isnan(a->score), bn = isnan(b->score);
  if (an || bn) return !an && bn;
  return a->score > b->score;
}

/* stable merge sort == Rust's slice::sort_by stability guarantee (index.rs:287) */
static void merge_sort(cxo_hit *v, cxo_hit *tmp, size_t n) {
  if (n < 2) return;
  if (n <= 16) {
    for (size_t i = 1; i < n; ++i) {
      cxo_hit x = v[i];
      size_t j = i;
      while (j > 0 && hit_before(&x, &v[j - 1])) {
        v[j] = v[j - 1];
        --j;
      }
      v[j] = x;
    }
    return;
  }
  size_t h = n / 2;
  merge_sort(v, tmp, h);
  merge_sort(v + h, tmp, n - h);
  size_t i = 0, j = h, o = 0;
  while (i < h && j < n) tmp[o++] = hit_before(&v[j], &v[i]) ? v[j++] : v[i++];
  while (i < h) tmp[o++] = v[i++];
  while (j < n) tmp[o++] = v[j++];
  memcpy(v, tmp, n * sizeof(cxo_hit));
}

/* vector/index.rs:259-294.  Returns all filtered hits sorted; caller truncates. */
static size_t scan_sorted(const cxo_index *ix, const float *q, size_t qlen, const cxo_filter *f,
                          cxo_hit **out) {
  cxo_hit *hits = (cxo_hit *)malloc((ix->n_live ? ix->n_live : 1) * sizeof(cxo_hit));
  size_t n = 0;
  for (size_t r = 0; r < ix->n_rows; ++r) {
    const cxo_row *R = &ix->rows[r];
    if (!R->live) continue;
    const float *row = ix->vecs + r * ix->dim;
    float d;
    if (ix->faithful_copy) { /* EmbeddingPoint(vec.clone()) per pair, index.rs:270 */
      float *cl = (float *)malloc((ix->dim ? ix->dim : 1) * sizeof(float));
      memcpy(cl, row, ix->dim * sizeof(float));
      d = cxo_distance(q, qlen, cl, ix->dim);
      free(cl);
    } else {
      d = cxo_distance(q, qlen, row, ix->dim);
    }
    if (f && !matches_filter(ix, R, f)) continue;
    hits[n].row = (uint32_t)r;
    hits[n].distance = d;
    hits[n].score = cxo_distance_to_similarity(d);
    ++n;
  }
  cxo_hit *tmp = (cxo_hit *)malloc((n ? n : 1) * sizeof(cxo_hit));
  merge_sort(hits, tmp, n);
  free(tmp);
  *out = hits;
  return n;
}

/* vector/index.rs:325-340 (index == None branch): empty -> 0 results. */
size_t cxo_search(const cxo_index *ix, const float *q, size_t qlen, size_t k, const cxo_filter *f,
                  uint8_t *out_ids, float *out_score, float *out_dist, uint32_t *out_rows) {
  if (ix->n_live == 0) return 0;
  cxo_hit *hits;
  size_t n = scan_sorted(ix, q, qlen, f, &hits);
  if (n > k) n = k;
  for (size_t i = 0; i < n; ++i) {
    if (out_ids) memcpy(out_ids + 16 * i, ix->rows[hits[i].row].id, 16);
    if (out_score) out_score[i] = hits[i].score;
    if (out_dist) out_dist[i] = hits[i].distance;
    if (out_rows) out_rows[i] = hits[i].row;
  }
  free(hits);
  return n;
}

/* vector/index.rs:376-388: search(q, max(len,1)) then keep score >= threshold.
 * Returns the full count; writes at most cap entries. */
size_t cxo_search_threshold(const cxo_index *ix, const float *q, size_t qlen, float threshold,
                            const cxo_filter *f, size_t cap, uint8_t *out_ids, float *out_score,
                            float *out_dist, uint32_t *out_rows) {
  if (ix->n_live == 0) return 0;
  cxo_hit *hits;
  size_t n = scan_sorted(ix, q, qlen, f, &hits);
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!(hits[i].score >= threshold)) continue;
    if (m < cap) {
      if (out_ids) memcpy(out_ids + 16 * m, ix->rows[hits[i].row].id, 16);
      if (out_score) out_score[m] = hits[i].score;
      if (out_dist) out_dist[m] = hits[i].distance;
      if (out_rows) out_rows[m] = hits[i].row;
    }
    ++m;
  }
  free(hits);
  return m;
}

/* vector/index.rs:390-410: rayon par_iter over queries == omp parallel for.
 * Outputs are [B][k]; out_n[b] = results for query b.  The reference collects
 * into a HashMap keyed by the query's NodeId; the host wrappers do that. */
int cxo_search_batch(const cxo_index *ix, const float *Q, size_t B, size_t qlen, size_t k,
                     const cxo_filter *f, uint8_t *out_ids, float *out_score, float *out_dist,
                     uint32_t *out_rows, uint64_t *out_n, int n_threads) {
  (void)n_threads;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (long b = 0; b < (long)B; ++b) {
    out_n[b] = cxo_search(ix, Q + (size_t)b * qlen, qlen, k, f, out_ids ? out_ids + 16 * k * b : NULL,
                          out_score ? out_score + k * b : NULL, out_dist ? out_dist + k * b : NULL,
                          out_rows ? out_rows + k * b : NULL);
  }
  return CXO_OK;
}

int cxo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* save / load: bincode 1.3 default config (fixint little endian) of          */
/* (&HashMap<Uuid,Vec<f32>>, &HashMap<Uuid,NodeMetadata>, usize)              */
/* vector/index.rs:437-472.  Uuid serialises as bytes (u64 len = 16 + 16 B),  */
/* NodeKind / source_agent as strings (u64 len + utf8).                       */

static void w64(FILE *fp, uint64_t v) { fwrite(&v, 8, 1, fp); }

int cxo_save(const cxo_index *ix, const char *path) {
  FILE *fp = fopen(path, "wb");
  if (!fp) return CXO_ERR_IO;
  w64(fp, ix->n_live);
  for (size_t r = 0; r < ix->n_rows; ++r) {
    if (!ix->rows[r].live) continue;
    w64(fp, 16);
    fwrite(ix->rows[r].id, 16, 1, fp);
    w64(fp, ix->dim);
    fwrite(ix->vecs + r * ix->dim, sizeof(float), ix->dim, fp);
  }
  size_t n_meta = ix->n_orphan;
  for (size_t r = 0; r < ix->n_rows; ++r)
    if (ix->rows[r].live && ix->rows[r].has_meta) ++n_meta;
  w64(fp, n_meta);
  for (size_t r = 0; r < ix->n_rows; ++r) {
    const cxo_row *R = &ix->rows[r];
    if (!R->live || !R->has_meta) continue;
    w64(fp, 16);
    fwrite(R->id, 16, 1, fp);
    w64(fp, strlen(ix->strs[R->kind]));
    fwrite(ix->strs[R->kind], 1, strlen(ix->strs[R->kind]), fp);
    w64(fp, strlen(ix->strs[R->agent]));
    fwrite(ix->strs[R->agent], 1, strlen(ix->strs[R->agent]), fp);
  }
  for (size_t o = 0; o < ix->n_orphan; ++o) {
    w64(fp, 16);
    fwrite(ix->orphan_ids + 16 * o, 16, 1, fp);
    const char *k = ix->strs[ix->orphan_kind[o]], *a = ix->strs[ix->orphan_agent[o]];
    w64(fp, strlen(k));
    fwrite(k, 1, strlen(k), fp);
    w64(fp, strlen(a));
    fwrite(a, 1, strlen(a), fp);
  }
  w64(fp, ix->dim);
  int bad = ferror(fp);
  fclose(fp);
  return bad ? CXO_ERR_IO : CXO_OK;
}

static int r64(FILE *fp, uint64_t *v) { return fread(v, 8, 1, fp) == 1; }

static char *rstr(FILE *fp) {
  uint64_t n;
  if (!r64(fp, &n) || n > (1u << 20)) return NULL;
  char *s = (char *)malloc(n + 1);
  if (n && fread(s, 1, n, fp) != n) {
    free(s);
    return NULL;
  }
  s[n] = 0;
  return s;
}

cxo_index *cxo_load(const char *path) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return NULL;
  /* dimension is the trailing usize */
  uint64_t dim = 0;
  if (fseek(fp, -8, SEEK_END) != 0 || !r64(fp, &dim)) {
    fclose(fp);
    return NULL;
  }
  fseek(fp, 0, SEEK_SET);
  cxo_index *ix = cxo_create((size_t)dim);
  uint64_t n, l;
  uint8_t id[16];
  if (!r64(fp, &n)) goto bad;
  for (uint64_t i = 0; i < n; ++i) {
    if (!r64(fp, &l) || l != 16 || fread(id, 16, 1, fp) != 1) goto bad;
    if (!r64(fp, &l)) goto bad;
    float *v = (float *)malloc((l ? l : 1) * sizeof(float));
    if (l && fread(v, sizeof(float), l, fp) != l) {
      free(v);
      goto bad;
    }
    /* load() does not re-validate lengths (index.rs:454-466); rows whose
     * length differs from `dimension` cannot be represented here */
    int rc = cxo_insert(ix, id, v, (size_t)l);
    free(v);
    if (rc != CXO_OK) goto bad;
  }
  if (!r64(fp, &n)) goto bad;
  for (uint64_t i = 0; i < n; ++i) {
    if (!r64(fp, &l) || l != 16 || fread(id, 16, 1, fp) != 1) goto bad;
    char *k = rstr(fp), *a = rstr(fp);
    if (!k || !a) {
      free(k);
      free(a);
      goto bad;
    }
    cxo_set_metadata(ix, id, k, a);
    free(k);
    free(a);
  }
  fclose(fp);
  return ix;
bad:
  fclose(fp);
  cxo_destroy(ix);
  return NULL;
}
