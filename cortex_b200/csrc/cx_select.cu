// cx_select.cu -- K5 + K3: order the nominated candidates, rescore the head with
// reference arithmetic, verify, emit.
//
// One CTA per query (small shared-memory footprint so several queries share an SM).
//   1. the query's merged candidate list (approximate keys, global memory) is cut down
//      to the keys that can still reach the top KP: everything at or above the pass's
//      final cut-off, and -- when that is still more than a few hundred -- at or above
//      the KP-th best score found by a radix select.  The survivors are sorted
//      (block-wide bitonic sort).  Meanwhile one thread computes the query norm in
//      reference order (index.rs:173)
//   2. the KS = min(#survivors, KP) best are rescored exactly: rows are staged into
//      shared memory with coalesced loads, then one thread per row runs the
//      reference's strict left-to-right fp32 fold (index.rs:172) -- so scores and
//      distances that leave here are bit-identical to the reference's
//   3. the rescored rows are ordered by (score desc, NaN last, row asc) -- the stable
//      descending sort of index.rs:287-292 with row order as the (reference-
//      unspecified) tie order -- and the first k are emitted
//   4. verification: U = best approximate cosine among everything NOT rescored (the
//      next key of the list, the radix cut and the pass's global cut-off).  The result
//      is exact if sim_k > U + eps, where eps bounds |approximate - reference| for the
//      pass that nominated the candidates, and score_k > 0 (below that the clamp of
//      index.rs:255 creates ties the approximate order cannot see).  Otherwise
//      ok[q] = 0 and the host reruns the query on a tighter path.
#include "cx_kernels.h"

namespace cx {

constexpr int SEL_MAX_KS = 512;            // rows that can be rescored for one query (head + the +-eps band)
constexpr uint32_t SEL_STAGE = 2048;      // radix-select staging words
constexpr double SCORE_QUANTUM_MARGIN = 1.1920928955078125e-7;  // 2^-23: twice the score quantum below 0.5

// Two launch shapes.  SMALL (keep count <= 32, i.e. k <= 16: interactive searches, search_batch top-10):
// 256 threads, rows rescored 32 at a time in chunks of 128 dimensions, ~31 KB of shared memory -- seven
// queries share an SM, so a batch of 1024 is resident at once and the kernel takes about as long as ONE
// query's chain of dependent steps.  LARGE (k = 100 auto-link scans, wide retries): 512 threads, 128 rows at
// a time.
template <bool SMALL>
struct SelShape {
  static constexpr int T = SMALL ? 256 : 512;
  static constexpr int PAR = SMALL ? 32 : 128;                        // rows rescored side by side
  static constexpr uint32_t STAGE_FLOATS = SMALL ? 32 * 129 : 128 * 65;  // [rows][W + 1]
};
__host__ __device__ inline bool select_small(uint32_t KP) { return KP <= 32; }

struct SelectParams {
  StoreView st;
  const float* Q;
  float* qnorm;    // written here (reference-order norm), read by the exact path
  float* rqnorm;
  uint32_t ldq, qlen;
  const uint64_t* keys;  // [nq][cap]
  uint32_t* cnt;         // [nq]   (re-zeroed on exit)
  uint64_t* gtau;        // [nq]   (re-zeroed on exit)
  uint32_t cap, KP;
  ResultView rv;         // pointers already offset to the first query of this launch
  float eps;
  int scale_by_rqn;      // pass keys are cosine * |q| (streaming pass) rather than cosine
  const uint32_t* qmap;  // optional: block i serves query qmap[i] (candidate list slot i)
};

struct SelectLayout {
  size_t keys, rstage, q, stage, e, total;
  uint32_t n_keep_max, keys_cap;
};

__host__ __device__ inline SelectLayout select_layout(uint32_t ld, uint32_t KP) {
  SelectLayout L;
  L.n_keep_max = 4 * KP < (uint32_t)SEL_MAX_KS ? 4 * KP : (uint32_t)SEL_MAX_KS;  // head + band
  L.keys_cap = 256;  // survivors incl. ties at the cut: a power of two (the bitonic fallback pads to one)
  while (L.keys_cap < 2 * L.n_keep_max) L.keys_cap <<= 1;
  size_t o = 0;
  L.keys = o;
  o += (size_t)L.keys_cap * 8;
  L.rstage = o;
  o += (size_t)SEL_STAGE * 4;
  L.q = o;
  o += (size_t)ld * 4;
  o = (o + 15) & ~(size_t)15;
  L.stage = o;
  o += (size_t)(select_small(KP) ? SelShape<true>::STAGE_FLOATS : SelShape<false>::STAGE_FLOATS) * 4;
  o = (o + 7) & ~(size_t)7;
  L.e = o;
  o += (size_t)L.n_keep_max * (8 + 4 + 4 + 4);
  L.total = o;
  return L;
}

// why queries fail verification (diagnostics, per device since process start): [0] candidate list overflowed
// or was truncated, [1] the margin test failed (near-ties around rank k denser than the rescored band),
// [2] everything else (too few candidates, degenerate query norm, NaN / zero k-th score)
__device__ unsigned long long g_select_why[4];

void select_why_read(uint64_t out[3]) {
  unsigned long long v[4] = {0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(v, g_select_why, sizeof v) != cudaSuccess) (void)cudaGetLastError();
  out[0] = v[0];
  out[1] = v[1];
  out[2] = v[2];
}

template <bool SMALL>
__global__ void __launch_bounds__(SelShape<SMALL>::T, SMALL ? 5 : 2) select_rescore_kernel(const SelectParams p) {
  constexpr int SEL_THREADS = SelShape<SMALL>::T;
  constexpr int SEL_PAR = SelShape<SMALL>::PAR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SelectLayout L = select_layout(p.st.ld, p.KP);
  const uint32_t SEL_K2 = L.keys_cap;
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw + L.keys);
  uint32_t* rstage = reinterpret_cast<uint32_t*>(smem_raw + L.rstage);
  float* q_s = reinterpret_cast<float*>(smem_raw + L.q);
  float* stage = reinterpret_cast<float*>(smem_raw + L.stage);
  uint64_t* ekey = reinterpret_cast<uint64_t*>(smem_raw + L.e);
  float* esim = reinterpret_cast<float*>(ekey + L.n_keep_max);
  float* edist = esim + L.n_keep_max;
  float* escore = edist + L.n_keep_max;
  __shared__ float s_na, s_simk, s_scorek;
  __shared__ uint32_t scratch[260];
  __shared__ uint32_t s_m;

  const uint32_t tid = threadIdx.x;
  const uint32_t q = p.qmap ? p.qmap[blockIdx.x] : blockIdx.x;  // query index (Q, state, results)
  const uint32_t ld = p.st.ld, dim = p.st.dim;
  const uint32_t n_app = p.cnt[q];
  const unsigned long long gt = p.gtau[q];  // final cut-off: nothing below it can reach the top KP
  const bool overflow = n_app > p.cap;
  const uint32_t n_src = min(n_app, p.cap);
  const uint64_t* src = p.keys + (size_t)blockIdx.x * p.cap;
  if (tid == 0) {
    s_simk = 0.0f;
    s_scorek = 0.0f;
    s_m = 0;
  }
  for (uint32_t d = tid; d < ld; d += SEL_THREADS) q_s[d] = d < p.qlen ? p.Q[(size_t)q * p.ldq + d] : 0.0f;
  __syncthreads();
  if (tid == SEL_THREADS - 1) {  // query norm, reference order (overlaps the list reads below)
    float acc = 0.0f;
    const uint32_t n = p.qlen < ld ? p.qlen : ld;
    uint32_t d = 0;
    for (; d + 8 <= n; d += 8) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = q_s[d + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = ref_fold(acc, x[j], x[j]);
    }
    for (; d < n; ++d) acc = ref_fold(acc, q_s[d], q_s[d]);
    const float na = __fsqrt_rn(acc);
    s_na = na;
    p.qnorm[q] = na;
    p.rqnorm[q] = __frcp_rn(na);
  }

  // ---- 1. survivors -> keys[0..M), ordered ---------------------------------------------
  // Only the best n_keep candidates can matter: the KS = KP that are rescored plus up to three times as many
  // again for the +-eps band around the k-th result (step 3).  A radix select finds that cut;
  // what it leaves behind is bounded by `cut` and enters U.
  unsigned long long U = gt;
  uint32_t cut = 0;  // radix cut (score ord); 0 = none
  const uint32_t n_keep = min(L.n_keep_max, n_src);
  if (n_src > n_keep) {
    auto get = [&](uint32_t i) {
      const uint64_t key = src[i];
      return key >= gt ? key_ord(key) : 0u;
    };
    cut = block_kth_largest(get, n_src, n_keep, scratch, rstage, SEL_STAGE, tid, SEL_THREADS);
    const unsigned long long left = (unsigned long long)cut << 32;  // what is left behind scores below `cut`
    if (left > U) U = left;
  }
  for (uint32_t i0 = 0; i0 < n_src; i0 += SEL_THREADS) {  // one counter update per warp
    const uint32_t i = i0 + tid;
    const uint64_t key = i < n_src ? src[i] : 0ull;
    const bool keep = i < n_src && key >= gt && key_ord(key) >= cut;
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    uint32_t base = 0;
    if ((tid & 31u) == 0 && bal) base = atomicAdd(&s_m, (uint32_t)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
      const uint32_t pos = base + (uint32_t)__popc(bal & ((1u << (tid & 31u)) - 1u));
      if (pos < SEL_K2) keys[pos] = key;
    }
  }
  __syncthreads();
  const uint32_t M = min(s_m, SEL_K2);
  const bool truncated = s_m > SEL_K2;
  if (M <= (uint32_t)SEL_THREADS) {
    // rank counting (keys are distinct): no barriers inside, every thread reads the same key at a time
    uint64_t mine = 0ull;
    uint32_t rank = 0;
    if (tid < M) {
      mine = keys[tid];
      for (uint32_t j = 0; j < M; ++j) rank += keys[j] > mine;
    }
    __syncthreads();
    if (tid < M) keys[rank] = mine;
    __syncthreads();
  } else {
    uint32_t NK = 32;
    while (NK < M) NK <<= 1;
    for (uint32_t i = M + tid; i < NK; i += SEL_THREADS) keys[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(keys, NK, tid, SEL_THREADS, [] { __syncthreads(); });
    __syncthreads();
  }
  const float na = s_na;
  uint32_t KS = min(min(M, p.KP), L.n_keep_max);

  // ---- 2. exact rescore of keys[lo, hi) ----------------------------------------------------------------
  // Up to SEL_PAR rows at a time, ALL of them in parallel: the rows are staged into shared memory in
  // column chunks of W dimensions ([rows][W]; LARGE: W = 256 / 128 / 64 for <= 32 / 64 / 128 rows, SMALL: 32 rows
  // x 128), one thread per
  // row carries that row's strict left-to-right fold (index.rs:172) across the chunks, and the next chunk is
  // already on its way from HBM (in registers) while the current one is folded.  The order of operations
  // inside a row is exactly the reference's; only independent rows run side by side.
  auto rescore = [&](uint32_t lo, uint32_t hi) {
    const uint32_t n_dims = p.qlen < dim ? p.qlen : dim;
    for (uint32_t base = lo; base < hi; base += SEL_PAR) {
      const uint32_t nb = min((uint32_t)SEL_PAR, hi - base);
      const uint32_t W = SMALL ? 128u : (nb <= 32 ? 256u : nb <= 64 ? 128u : 64u);   // floats per row per chunk
      const uint32_t W4 = W >> 2, sstride = W + 1;
      const uint32_t n_chunks = (ld + W - 1) / W;
      // element e of a chunk = (row e / W4, float4 e % W4); thread tid owns e = tid + i * SEL_THREADS, i < 4
      float4 nxt[4];
      auto fetch = [&](uint32_t c) {
        const uint32_t w4 = min(W4, (ld - c * W) >> 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t e = tid + i * SEL_THREADS, r = e / W4, f = e % W4;
          if (r < nb && f < w4) {
            const uint32_t row = key_row(keys[base + r]);
            nxt[i] = __ldg(reinterpret_cast<const float4*>(p.st.E + (size_t)row * ld + c * W) + f);
          }
        }
      };
      float dot = 0.0f;
      fetch(0);
      for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint32_t w4 = min(W4, (ld - c * W) >> 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t e = tid + i * SEL_THREADS, r = e / W4, f = e % W4;
          if (r < nb && f < w4) {
            float* d = stage + r * sstride + 4 * f;
            d[0] = nxt[i].x;
            d[1] = nxt[i].y;
            d[2] = nxt[i].z;
            d[3] = nxt[i].w;
          }
        }
        __syncthreads();
        if (c + 1 < n_chunks) fetch(c + 1);  // in flight while this chunk is folded
        if (tid < nb) {
          const float* r = stage + tid * sstride;
          const float* qq = q_s + c * W;
          const uint32_t d0 = c * W;
          const uint32_t n = n_dims > d0 ? min(W, n_dims - d0) : 0u;
          uint32_t d = 0;
          for (; d + 8 <= n; d += 8) {
            float a[8], b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              a[j] = qq[d + j];
              b[j] = r[d + j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) dot = ref_fold(dot, a[j], b[j]);
          }
          for (; d < n; ++d) dot = ref_fold(dot, qq[d], r[d]);
        }
        __syncthreads();
      }
      if (tid < nb) {
        const uint32_t row = key_row(keys[base + tid]);
        const float nbm = __ldg(p.st.norm + row);
        const float sim = __fdiv_rn(dot, __fmul_rn(na, nbm));
        const float dist = __fsub_rn(1.0f, sim);
        const float sc = ref_score_from_distance(dist);
        ekey[base + tid] = make_key(ord_from_score(sc), row);
        esim[base + tid] = sim;
        edist[base + tid] = dist;
        escore[base + tid] = sc;
      }
      __syncthreads();
    }
  };
  // rank the first `n_res` rescored rows by exact key (all distinct: the row is part of the
  // key); remember the k-th; optionally emit the first k
  const uint32_t k = p.rv.k;
  auto rank_rows = [&](uint32_t n_res, bool emit) {
    const uint32_t n_out = min(k, n_res);
    if (tid < n_res) {
      const uint64_t mine = ekey[tid];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < n_res; ++j) rank += ekey[j] > mine;
      if (emit && rank < n_out) {
        const size_t o = (size_t)q * k + rank;
        const uint32_t row = key_row(mine);
        p.rv.rows[o] = row;
        p.rv.score[o] = escore[tid];
        p.rv.dist[o] = edist[tid];
        if (p.rv.ids)
          *reinterpret_cast<uint4*>(p.rv.ids + o * 16) =
              *reinterpret_cast<const uint4*>(p.st.ids + (size_t)row * 16);
      }
      if (rank + 1 == k) {
        s_simk = esim[tid];
        s_scorek = escore[tid];
      }
    }
    __syncthreads();
  };

  rescore(0, KS);
  // ---- 3. widen to the whole +-eps band when the list covers it ---------------------------
  // Every candidate whose approximate cosine is within eps of the current k-th exact cosine
  // could still belong to the top k.  The list holds all rows down to its cut-off, so if the
  // band lies above the cut-off, rescoring the band makes the answer exact without a re-scan.
  if (KS >= k && KS < M && p.eps > 0.0f) {
    rank_rows(KS, false);
    const float simk = s_simk;
    if (simk == simk) {
      float band = simk - p.eps - 2.0f * (float)SCORE_QUANTUM_MARGIN;  // cosine units; covers the verify margin
      if (p.scale_by_rqn) band = band * na;      // streaming-pass keys are cosine * |q|
      uint32_t KS1 = KS;
      while (KS1 < M && KS1 < L.n_keep_max && float_from_ord(key_ord(keys[KS1])) >= band) ++KS1;
      if (KS1 > KS) {
        rescore(KS, KS1);
        KS = KS1;
      }
    }
  }
  if (M > KS && keys[KS] > U) U = keys[KS];
  rank_rows(KS, true);

  // ---- 4. verify ---------------------------------------------------------------------
  if (tid == 0) {
    p.rv.n[q] = min(k, KS);
    bool ok = KS >= k && !overflow && !truncated && na >= NORM_REGULAR_MIN && na <= NORM_REGULAR_MAX;
    int why = (overflow || truncated) ? 0 : 2;
    if (ok) {
      const float sk = s_scorek, simk = s_simk;
      if (sk != sk) ok = false;
      else if (U != 0ull) {
        why = sk > 0.0f ? 1 : 2;
        float u = float_from_ord(key_ord(U));
        if (p.scale_by_rqn) u = u * __frcp_rn(na);
        // every row that was not rescored has a reference cosine <= u + eps.  Its SCORE must be strictly
        // below score_k (an equal score with a lower row would sort ahead): 1 - (1 - sim) rounds sim to a
        // multiple of 2^-24 at worst (index.rs:177,255), so two cosines more than 2^-23 apart can never
        // produce the same score.  Summed in double so that the margin itself is not rounded away.
        ok = (sk > 0.0f) && ((double)simk > (double)u + (double)p.eps + SCORE_QUANTUM_MARGIN);
      }
    }
    p.rv.ok[q] = ok ? 1u : 0u;
    if (!ok) atomicAdd(&g_select_why[why], 1ull);
    p.cnt[q] = 0;   // leave the workspace ready for the next pass
    p.gtau[q] = 0ull;
  }
}

// ---------------------------------------------------------------------------------
// Threshold scans (search_threshold, index.rs:376-388; the dedup scanner's per-node search,
// linker/dedup.rs:84-86): the pass nominated every row whose approximate cosine reaches
// threshold - eps.  Here ALL nominees are rescored with reference arithmetic, the exact test
// `score >= threshold` (index.rs:385) is applied, and the survivors come out ordered by
// (score desc, row asc).  Because the nominees are a superset of the qualifying rows, the
// result is exact with no verification step; the only failure is a list that overflowed
// (ok = 0 -> the host reruns that query on the exact path).
constexpr int THR_THREADS = 512;
constexpr int THR_PAR = 128;                       // nominees rescored side by side
constexpr uint32_t THR_STAGE_FLOATS = 128 * 65;    // [rows][W + 1]: 128 x 64, 64 x 128 or 32 x 256 floats per chunk
constexpr uint32_t THR_MAX = 2048;  // survivors that can be ordered in shared memory

struct ThresholdParams {
  StoreView st;
  const float* Q;
  const float* qnorm;    // reference-order norms (launch_prepare_queries)
  uint32_t ldq, qlen;
  const uint64_t* keys;  // [nq][cap] nominees (approximate keys; only the row is used)
  uint32_t* cnt;         // [nq] (re-zeroed on exit)
  uint64_t* gtau;        // [nq] (re-zeroed on exit)
  uint32_t cap;
  float threshold;
  const uint32_t* self_rows;  // optional [nq]: row to skip (the query's own row), 0xFFFFFFFF = none
  uint32_t upper_only;        // 1: keep only rows above self_rows[q] (unordered pairs, each once)
  const uint64_t* q_seq;      // optional [nq] + row_seq [rows]: keep only rows with row_seq[row] > q_seq[q]
  const uint64_t* row_seq;    //   (pair scans of a multi-device index: global insertion numbers)
  ResultView rv;         // k = output capacity per query; n = written; ok
  uint32_t* total;       // [nq] how many qualify
};

struct ThresholdLayout {
  size_t q, stage, ekey, edist, escore, idx, total;
};
__host__ __device__ inline ThresholdLayout threshold_layout(uint32_t ld) {
  ThresholdLayout L;
  size_t o = 0;
  L.q = o;
  o += (size_t)ld * 4;
  o = (o + 15) & ~(size_t)15;
  L.stage = o;
  o += (size_t)THR_STAGE_FLOATS * 4;
  o = (o + 7) & ~(size_t)7;
  L.ekey = o;
  o += (size_t)THR_MAX * 8;
  L.edist = o;
  o += (size_t)THR_MAX * 4;
  L.escore = o;
  o += (size_t)THR_MAX * 4;
  L.idx = o;
  o += (size_t)THR_MAX * 2;
  L.total = o;
  return L;
}

__global__ void __launch_bounds__(THR_THREADS) threshold_rescore_kernel(const ThresholdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ThresholdLayout L = threshold_layout(p.st.ld);
  float* q_s = reinterpret_cast<float*>(smem_raw + L.q);
  float* stage = reinterpret_cast<float*>(smem_raw + L.stage);
  uint64_t* ekey = reinterpret_cast<uint64_t*>(smem_raw + L.ekey);
  float* edist = reinterpret_cast<float*>(smem_raw + L.edist);
  float* escore = reinterpret_cast<float*>(smem_raw + L.escore);
  uint16_t* eidx = reinterpret_cast<uint16_t*>(smem_raw + L.idx);
  __shared__ uint32_t s_m;

  const uint32_t tid = threadIdx.x, q = blockIdx.x;
  const uint32_t ld = p.st.ld, dim = p.st.dim;
  const uint32_t n_app = p.cnt[q];
  const uint32_t n_src = min(n_app, p.cap);
  const uint64_t* src = p.keys + (size_t)q * p.cap;
  const uint32_t self = p.self_rows ? p.self_rows[q] : 0xFFFFFFFFu;
  const bool by_seq = p.q_seq != nullptr && p.row_seq != nullptr;
  const uint64_t my_seq = by_seq ? p.q_seq[q] : 0ull;
  if (tid == 0) s_m = 0;
  for (uint32_t d = tid; d < ld; d += THR_THREADS) q_s[d] = d < p.qlen ? p.Q[(size_t)q * p.ldq + d] : 0.0f;
  __syncthreads();
  const float na = p.qnorm[q];
  // All nominees are rescored, THR_PAR rows side by side: the rows are staged in column chunks of W dimensions,
  // one thread per row carries its strict left-to-right fold (index.rs:172) across the chunks while the next
  // chunk is already in flight (the scheme of the select kernel above).
  const uint32_t n_dims = p.qlen < dim ? p.qlen : dim;
  for (uint32_t base = 0; base < n_src; base += THR_PAR) {
    const uint32_t nb = min((uint32_t)THR_PAR, n_src - base);
    const uint32_t W = nb <= 32 ? 256u : nb <= 64 ? 128u : 64u;
    const uint32_t W4 = W >> 2, sstride = W + 1;
    const uint32_t n_chunks = (ld + W - 1) / W;
    float4 nxt[4];
    auto fetch = [&](uint32_t c) {
      const uint32_t w4 = min(W4, (ld - c * W) >> 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t e = tid + i * THR_THREADS, r = e / W4, f = e % W4;
        if (r < nb && f < w4) {
          const uint32_t row = key_row(src[base + r]);
          nxt[i] = __ldg(reinterpret_cast<const float4*>(p.st.E + (size_t)row * ld + c * W) + f);
        }
      }
    };
    float dot = 0.0f;
    fetch(0);
    for (uint32_t c = 0; c < n_chunks; ++c) {
      const uint32_t w4 = min(W4, (ld - c * W) >> 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t e = tid + i * THR_THREADS, r = e / W4, f = e % W4;
        if (r < nb && f < w4) {
          float* d = stage + r * sstride + 4 * f;
          d[0] = nxt[i].x;
          d[1] = nxt[i].y;
          d[2] = nxt[i].z;
          d[3] = nxt[i].w;
        }
      }
      __syncthreads();
      if (c + 1 < n_chunks) fetch(c + 1);
      if (tid < nb) {
        const float* r = stage + tid * sstride;
        const float* qq = q_s + c * W;
        const uint32_t d0 = c * W;
        const uint32_t n = n_dims > d0 ? min(W, n_dims - d0) : 0u;
        uint32_t d = 0;
        for (; d + 8 <= n; d += 8) {
          float a[8], b[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a[j] = qq[d + j];
            b[j] = r[d + j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) dot = ref_fold(dot, a[j], b[j]);
        }
        for (; d < n; ++d) dot = ref_fold(dot, qq[d], r[d]);
      }
      __syncthreads();
    }
    if (tid < nb) {
      const uint32_t row = key_row(src[base + tid]);
      const bool wanted = by_seq ? __ldg(p.row_seq + row) > my_seq : (row != self && !(p.upper_only && row < self));
      if (wanted) {
        const float nbm = __ldg(p.st.norm + row);
        const float sim = __fdiv_rn(dot, __fmul_rn(na, nbm));
        const float dist = __fsub_rn(1.0f, sim);
        const float sc = ref_score_from_distance(dist);
        if (sc >= p.threshold) {  // index.rs:385 (false for NaN)
          const uint32_t pos = atomicAdd(&s_m, 1u);
          if (pos < THR_MAX) {
            ekey[pos] = make_key(ord_from_score(sc), row);
            edist[pos] = dist;
            escore[pos] = sc;
            eidx[pos] = (uint16_t)pos;
          }
        }
      }
    }
    __syncthreads();
  }

  const uint32_t M = s_m;
  // a query whose norm under- / overflowed is not covered by the nominating pass's error bound: exact
  // path.  (In a pair scan the queries are rows of the index, all regular or all-zero there: an all-zero
  // row scores NaN against everything and simply has no partners.)
  const bool q_regular = (na >= NORM_REGULAR_MIN && na <= NORM_REGULAR_MAX) || p.self_rows != nullptr || by_seq;
  const bool ok = n_app <= p.cap && M <= THR_MAX && q_regular;
  if (ok && M) {
    // order (key, idx) by key descending: bitonic network over a power of two, padded with 0
    uint32_t NK = 2;
    while (NK < M) NK <<= 1;
    for (uint32_t i = M + tid; i < NK; i += THR_THREADS) {
      ekey[i] = 0ull;
      eidx[i] = 0;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= NK; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t i = tid; i < NK; i += THR_THREADS) {
          const uint32_t l = i ^ j;
          if (l > i) {
            const uint64_t x = ekey[i], y = ekey[l];
            const bool desc_block = (i & k) == 0;
            if (desc_block ? (x < y) : (x > y)) {
              ekey[i] = y;
              ekey[l] = x;
              const uint16_t t = eidx[i];
              eidx[i] = eidx[l];
              eidx[l] = t;
            }
          }
        }
        __syncthreads();
      }
    }
    const uint32_t n_out = min(M, p.rv.k);
    for (uint32_t i = tid; i < n_out; i += THR_THREADS) {
      const size_t o = (size_t)q * p.rv.k + i;
      const uint32_t row = key_row(ekey[i]);
      const uint32_t e = eidx[i];
      p.rv.rows[o] = row;
      p.rv.score[o] = escore[e];
      p.rv.dist[o] = edist[e];
      if (p.rv.ids)
        *reinterpret_cast<uint4*>(p.rv.ids + o * 16) = *reinterpret_cast<const uint4*>(p.st.ids + (size_t)row * 16);
    }
  }
  if (tid == 0) {
    p.rv.n[q] = ok ? min(M, p.rv.k) : 0u;
    p.rv.ok[q] = ok ? 1u : 0u;
    if (p.total) p.total[q] = ok ? M : 0u;
    p.cnt[q] = 0;
    p.gtau[q] = 0ull;
  }
}

cudaError_t launch_threshold_rescore(const StoreView& st, const QueryView& qv, uint32_t q0, uint32_t nq,
                                     const CandView& cv, const ResultView& rv, uint32_t* total, float threshold,
                                     const uint32_t* self_rows, bool upper_only, cudaStream_t s,
                                     const uint64_t* q_seq, const uint64_t* row_seq) {
  if (!nq) return cudaSuccess;
  ThresholdParams p;
  p.q_seq = q_seq ? q_seq + q0 : nullptr;
  p.row_seq = row_seq;
  p.st = st;
  p.Q = qv.Q + (size_t)q0 * qv.ldq;
  p.qnorm = qv.qnorm + q0;
  p.ldq = qv.ldq;
  p.qlen = qv.qlen;
  p.keys = cv.keys + (size_t)(q0 - cv.q_base) * cv.cap;
  p.cnt = cv.cnt + q0;
  p.gtau = cv.gtau + q0;
  p.cap = cv.cap;
  p.threshold = threshold;
  p.self_rows = self_rows ? self_rows + q0 : nullptr;
  p.upper_only = upper_only ? 1u : 0u;
  p.rv = rv;
  p.rv.rows += (size_t)q0 * rv.k;
  p.rv.score += (size_t)q0 * rv.k;
  p.rv.dist += (size_t)q0 * rv.k;
  if (p.rv.ids) p.rv.ids += (size_t)q0 * rv.k * 16;
  p.rv.n += q0;
  p.rv.ok += q0;
  p.total = total ? total + q0 : nullptr;
  const size_t smem = threshold_layout(st.ld).total;
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = raise_dynamic_smem<threshold_rescore_kernel>(smem);
  if (e != cudaSuccess) return e;
  threshold_rescore_kernel<<<nq, THR_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}
size_t threshold_rescore_smem(uint32_t ld) { return threshold_layout(ld).total; }

size_t select_smem(uint32_t cap, uint32_t ld) {
  (void)cap;
  return select_layout(ld, 128).total + 2048;  // the larger of the two launch shapes
}

cudaError_t launch_select_rescore(const StoreView& st, const QueryView& qv, uint32_t q0, uint32_t nq,
                                  const CandView& cv, const ResultView& rv, float eps_cos, int scale_by_rqn,
                                  cudaStream_t s, const uint32_t* qmap) {
  if (!nq) return cudaSuccess;
  if (qmap && q0 != 0) return cudaErrorInvalidValue;  // a query map addresses queries absolutely
  SelectParams p;
  p.qmap = qmap;
  p.st = st;
  p.Q = qv.Q + (size_t)q0 * qv.ldq;
  p.qnorm = const_cast<float*>(qv.qnorm) + q0;
  p.rqnorm = const_cast<float*>(qv.rqnorm) + q0;
  p.ldq = qv.ldq;
  p.qlen = qv.qlen;
  p.keys = cv.keys + (size_t)(q0 - cv.q_base) * cv.cap;
  p.cnt = cv.cnt + q0;
  p.gtau = cv.gtau + q0;
  p.cap = cv.cap;
  p.KP = cv.KP;
  p.rv = rv;
  p.rv.rows += (size_t)q0 * rv.k;
  p.rv.score += (size_t)q0 * rv.k;
  p.rv.dist += (size_t)q0 * rv.k;
  if (p.rv.ids) p.rv.ids += (size_t)q0 * rv.k * 16;
  p.rv.n += q0;
  p.rv.ok += q0;
  p.eps = eps_cos;
  p.scale_by_rqn = scale_by_rqn;
  const size_t smem = select_layout(st.ld, cv.KP).total;
  if (smem > 200 * 1024 || cv.KP > SEL_MAX_KS) return cudaErrorInvalidConfiguration;
  if (select_small(cv.KP)) {
    // with the largest shared-memory carve-out: the point of the small shape is several queries per SM
    cudaError_t e = raise_dynamic_smem<select_rescore_kernel<true>>(smem, true);
    if (e != cudaSuccess) return e;
    select_rescore_kernel<true><<<nq, SelShape<true>::T, smem, s>>>(p);
  } else {
    cudaError_t e = raise_dynamic_smem<select_rescore_kernel<false>>(smem, true);
    if (e != cudaSuccess) return e;
    select_rescore_kernel<false><<<nq, SelShape<false>::T, smem, s>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace cx
