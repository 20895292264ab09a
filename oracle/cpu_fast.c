/*
 * cpu_fast.c -- an OPTIMISED CPU scan, reported next to the faithful port as a courtesy baseline
 * (SURVEY.md 8d iii).  BASELINE INFRASTRUCTURE ONLY and NOT THE REFERENCE: it is what a CPU
 * implementation of the same job could do if it were rewritten the way the GPU path was --
 * row norms hoisted out of the pair loop, the corpus streamed once per block of queries, vectorised
 * fp32 dot products (reassociated, FMA allowed: compiled with -O3 -ffast-math for x86-64-v4 and -v3 in its own
 * translation unit), a k-entry heap per query instead of sorting N results, OpenMP over row blocks.
 * Scores are approximate cosines; ids are the exact top-k except where scores tie within rounding.
 * The reference does none of this (vector/index.rs:259-294: three scalar reductions, a heap allocation
 * and a memcpy per pair, an O(N log N) sort per query).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QB 16 /* queries that share one pass over a block of rows */

/* 16 floats: one AVX-512 register, or two AVX2 registers in the v3 build */
typedef float v16 __attribute__((vector_size(64), aligned(4)));
static inline float hsum(v16 v) {
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += v[i];
  return s;
}

typedef struct {
  float s;
  uint32_t r;
} hit_t;

/* min-heap on (score, then larger row first) so that the root is the weakest kept hit */
static inline int weaker(hit_t a, hit_t b) { return a.s < b.s || (a.s == b.s && a.r > b.r); }
static void heap_push(hit_t *h, size_t *n, size_t k, hit_t x) {
  if (*n < k) {
    size_t i = (*n)++;
    h[i] = x;
    while (i && weaker(h[i], h[(i - 1) / 2])) {
      hit_t t = h[i];
      h[i] = h[(i - 1) / 2];
      h[(i - 1) / 2] = t;
      i = (i - 1) / 2;
    }
    return;
  }
  if (!weaker(h[0], x)) return;
  h[0] = x;
  size_t i = 0;
  for (;;) {
    size_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < k && weaker(h[l], h[m])) m = l;
    if (r < k && weaker(h[r], h[m])) m = r;
    if (m == i) break;
    hit_t t = h[i];
    h[i] = h[m];
    h[m] = t;
    i = m;
  }
}
static int cmp_desc(const void *a, const void *b) {
  const hit_t *x = a, *y = b;
  if (x->s != y->s) return x->s < y->s ? 1 : -1;
  return x->r < y->r ? -1 : (x->r > y->r);
}

/* rnorm[n] = 1 / |row| */
void cxf_row_rnorms(const float *E, size_t n, size_t d, float *rnorm) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    const float *e = E + i * d;
    float acc = 0.f;
    for (size_t j = 0; j < d; ++j) acc += e[j] * e[j];
    rnorm[i] = 1.0f / sqrtf(acc);
  }
}

/* top-k (cosine, best first) of nq queries over n rows; out_rows/out_score are [nq][k].
 * The library travels to other hosts, so the Makefile builds it twice (x86-64-v4 = AVX-512 and
 * x86-64-v3 = AVX2 + FMA) and oracle/binding.py loads the one the host's CPU flags allow. */
void cxf_search_batch(const float *E, const float *rnorm, size_t n, size_t d, const float *Q, size_t nq, size_t k,
                      uint32_t *out_rows, float *out_score, int n_threads) {
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
  const int T = omp_get_max_threads();
#else
  const int T = 1;
#endif
  float *qn = malloc(nq * sizeof(float));
  for (size_t b = 0; b < nq; ++b) {
    float acc = 0.f;
    for (size_t j = 0; j < d; ++j) acc += Q[b * d + j] * Q[b * d + j];
    qn[b] = 1.0f / sqrtf(acc);
  }
  /* per thread, per query of the current block: a k-heap */
  hit_t *heaps = malloc((size_t)T * QB * k * sizeof(hit_t));
  size_t *cnt = malloc((size_t)T * QB * sizeof(size_t));
  float *Qp = malloc((size_t)QB * d * sizeof(float));
  for (size_t q0 = 0; q0 < nq; q0 += QB) {
    const size_t nb = nq - q0 < QB ? nq - q0 : QB;
    memset(cnt, 0, (size_t)T * QB * sizeof(size_t));
    memset(Qp, 0, (size_t)QB * d * sizeof(float));  /* the block of queries, zero padded to QB */
    memcpy(Qp, Q + q0 * d, nb * d * sizeof(float));
#pragma omp parallel
    {
#ifdef _OPENMP
      const int t = omp_get_thread_num();
#else
      const int t = 0;
#endif
      hit_t *hp = heaps + (size_t)t * QB * k;
      size_t *cp = cnt + (size_t)t * QB;
#pragma omp for schedule(static)
      for (size_t i = 0; i < n; ++i) {
        const float *e = E + i * d;
        float acc[QB];
        {
          /* one load of the row chunk feeds QB fused multiply-adds: QB vector accumulators stay in registers */
          v16 va[QB];
          for (size_t b = 0; b < QB; ++b) va[b] = (v16){0};
          size_t j = 0;
          for (; j + 16 <= d; j += 16) {
            const v16 ev = *(const v16 *)(e + j);
            for (size_t b = 0; b < QB; ++b) va[b] += ev * *(const v16 *)(Qp + b * d + j);
          }
          for (size_t b = 0; b < QB; ++b) {
            float a = hsum(va[b]);
            for (size_t jj = j; jj < d; ++jj) a += e[jj] * Qp[b * d + jj];
            acc[b] = a;
          }
        }
        const float rn = rnorm[i];
        for (size_t b = 0; b < nb; ++b) {
          hit_t x = {acc[b] * rn * qn[q0 + b], (uint32_t)i};
          if (x.s == x.s) heap_push(hp + b * k, cp + b, k, x);
        }
      }
    }
    /* merge the threads' heaps */
    for (size_t b = 0; b < nb; ++b) {
      size_t m = 0;
      hit_t *all = malloc((size_t)T * k * sizeof(hit_t));
      for (int t = 0; t < T; ++t) {
        const size_t c = cnt[(size_t)t * QB + b];
        memcpy(all + m, heaps + ((size_t)t * QB + b) * k, c * sizeof(hit_t));
        m += c;
      }
      qsort(all, m, sizeof(hit_t), cmp_desc);
      for (size_t j = 0; j < k; ++j) {
        out_rows[(q0 + b) * k + j] = j < m ? all[j].r : 0xFFFFFFFFu;
        out_score[(q0 + b) * k + j] = j < m ? all[j].s : NAN;
      }
      free(all);
    }
  }
  free(heaps);
  free(cnt);
  free(Qp);
  free(qn);
}
