// cx_stream.cu -- K1: small-batch streaming fp32 scan (HBM-bound).
//
// One persistent CTA per SM.  A producer thread streams 32-row tiles of the
// row-major embedding matrix (plus the 32 reciprocal norms that go with them)
// into a shared-memory ring with 1-D bulk async copies (cp.async.bulk -> UBLKCP,
// completion on mbarriers).  Sixteen consumer warps work as two groups that take
// alternate tiles; inside a group every warp owns four rows, lanes split the
// dimension (conflict-free LDS.128), NQ query dot products per row accumulate
// with FFMA, and a transposing warp reduction leaves every lane with one
// (row, query) score.
//
// Scores only nominate candidates.  Each CTA keeps a per-query candidate list in
// shared memory with a running cut-off tau (the KP-th best key seen so far, shared
// across CTAs through one global word per query), so after the first few tiles a
// score costs one compare.  Lists are merged, rescored with reference arithmetic
// and verified by cx_select.cu.  The matrix is read exactly once per pass:
//   algorithmic bytes per pass = n_rows * ld * 4   
//
// HALF variant (B <= 4 on an index that keeps the bf16 shadow): the same pipeline streams the NORMALISED bf16
// copy of the rows instead (half the bytes: n_rows * ld16 * 2 per pass), widens the elements in registers and
// needs no reciprocal norms.  Its scores carry the shadow's rounding error (2^-9 relative per row), which only
// widens the band cx_select.cu rescores exactly.
//
// Call shapes served (reference): search() at B=1 (api.rs:117-125,
// http/routes.rs:906-907, grpc/service.rs:673-675, gate/mod.rs:332) and small
// search_batch() groups (vector/index.rs:390-410).
#include "cx_kernels.h"

namespace cx {

constexpr int SC_GW = 8;                         // warps per consumer group
constexpr int SC_GROUPS = 2;                     // groups take alternate tiles
constexpr int SC_CW = SC_GW * SC_GROUPS;         // consumer warps
constexpr int SC_CT = SC_CW * 32;                // consumer threads
constexpr int SC_THREADS = SC_CT + 32;           // + 1 producer warp
constexpr int SC_R = 4;                          // rows per warp per tile
constexpr int SC_TILE_ROWS = SC_GW * SC_R;       // 32: rows of a tile of the fp32 variant (one group's share)
// The bf16 variant's tiles are 64 rows (the same bytes per stage as 32 fp32 rows): BOTH groups work on every
// tile, 32 rows each, so they reach the list checkpoints together instead of waiting a tile time for each other.
__host__ __device__ constexpr int sc_tile_rows(bool half) { return half ? SC_TILE_ROWS * SC_GROUPS : SC_TILE_ROWS; }
constexpr int SC_QUAD = 4;                       // tiles between list checks
constexpr uint32_t SC_EARLY_TILES = 16;          // bf16 variant: tiles checked at the short interval
constexpr int SC_KSLICE = 384;                   // floats of a row per stage (<= 48 KB stages)
constexpr int SC_MAX_STAGES = 8;                 // fp32 rows: 4 x 48 KB fit; bf16 rows: 8 x 24 KB
constexpr size_t SC_SMEM_LIMIT = 227 * 1024;

struct StreamParams {
  StoreView st;
  const float* Q;    // first query of this pass
  uint32_t ldq, nq_valid;
  DevFilter flt;
  uint64_t* keys;    // &cand.keys[q0][0]   merged candidate list per query, capacity cap
  uint32_t* cnt;     // &cand.cnt[q0]       its fill count        (zero before the pass)
  uint64_t* gtau;    // &cand.gtau[q0]      cross-CTA cut-off     (zero before the pass)
  uint32_t cap;
  const uint32_t* qmap;  // optional: pass-local query b is query qmap[b] (Q row, cnt, gtau); its list slot stays b
  uint32_t G, KP, C, n_tiles, stages, kslice, n_slices;
  uint32_t quad;         // bf16 variant: most tiles between two list checks (what the list capacity allows); the
                         // first SC_EARLY_TILES tiles are checked every second tile so the cut-off settles fast
  // threshold mode (search_threshold, index.rs:376-388): instead of a running top-KP cut-off every
  // row whose approximate cosine reaches thr_cos is nominated; full lists are flushed, not compacted
  uint32_t thr_mode;
  float thr_cos;         // threshold minus the pass's error bound, in cosine units
  const float* qnorm;    // |q| per query (keys are cosine * |q|)
};

template <int NV>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[NV], uint32_t lane) {
  // After the call lane L holds the warp-wide sum of v[L / (32/NV)].
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int off = 16 >> s;
    const int n = NV >> s;
    if (n > 1) {
      const int half = n >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        float send = upper ? v[i] : v[i + half];
        float keep = upper ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
  return v[0];
}

template <int NV>
__device__ __forceinline__ uint32_t transpose_index(uint32_t lane) {
  uint32_t idx = 0;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int n = NV >> s;
    if (n > 1) idx += ((lane >> (4 - s)) & 1u) * (uint32_t)(n >> 1);
  }
  return idx;
}

struct StreamLayout {
  size_t tiles, rn, q, list, tau, cnt, bars, total;
};

// ld: elements per stored row (and per query row in shared memory); esize: bytes per stored element
__host__ __device__ inline StreamLayout stream_layout(uint32_t ld, uint32_t nq, uint32_t C, uint32_t stages,
                                                      uint32_t kslice, uint32_t esize) {
  const uint32_t tile_rows = sc_tile_rows(esize == 2);
  StreamLayout L;
  size_t o = 0;
  L.tiles = o;
  o += (size_t)stages * tile_rows * kslice * esize;
  L.rn = o;
  o += (size_t)SC_MAX_STAGES * SC_TILE_ROWS * 4;
  L.q = o;
  o += (size_t)nq * ld * 4;
  o = (o + 15) & ~(size_t)15;
  L.list = o;
  o += (size_t)nq * 2 * C * 8;  // double buffered
  L.tau = o;
  o += (size_t)nq * 8;
  L.cnt = o;
  o += (size_t)(2 * nq + 4) * 4;  // cnt[nq], cur[nq], flag
  o = (o + 7) & ~(size_t)7;
  L.bars = o;
  o += (size_t)2 * SC_MAX_STAGES * 8;
  L.total = o;
  return L;
}

// KC > 0: the slice is exactly KC chunks of 128 floats (fully unrolled inner loop)
// HALF: rows come from the normalised bf16 shadow (st.E16, ld16 elements per row)
template <int NQ, int KC, bool HALF>
__global__ void __launch_bounds__(SC_THREADS, 1) stream_scan_kernel(const StreamParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t ES = HALF ? 2u : 4u;  // bytes per stored element
  constexpr uint32_t TR = sc_tile_rows(HALF);   // rows per tile
  const uint32_t QUAD = HALF ? p.quad : (uint32_t)SC_QUAD;  // tiles between list checks
  const uint32_t ld = HALF ? p.st.ld16 : p.st.ld;
  const StreamLayout L = stream_layout(ld, NQ, p.C, p.stages, p.kslice, ES);
  unsigned char* tiles = smem_raw + L.tiles;
  float* rn_s = reinterpret_cast<float*>(smem_raw + L.rn);
  float* q_s = reinterpret_cast<float*>(smem_raw + L.q);
  uint64_t* list_s = reinterpret_cast<uint64_t*>(smem_raw + L.list);
  uint64_t* tau_s = reinterpret_cast<uint64_t*>(smem_raw + L.tau);
  uint32_t* cnt_s = reinterpret_cast<uint32_t*>(smem_raw + L.cnt);
  uint32_t* cur_s = cnt_s + NQ;
  uint32_t* flag_s = cnt_s + 2 * NQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ld4 = ld >> 2, S = p.stages;
  const uint32_t tile_bytes = TR * p.kslice * ES;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + SC_MAX_STAGES);

  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, HALF ? SC_CW : SC_GW);  // consumer warps that read a stage
    }
    fence_mbar_init();
  }
  for (uint32_t i = tid; i < NQ * ld; i += SC_THREADS) {
    uint32_t b = i / ld, d = i % ld;
    q_s[i] = (b < p.nq_valid && d < p.ldq) ? p.Q[(size_t)(p.qmap ? p.qmap[b] : b) * p.ldq + d] : 0.0f;
  }
  for (uint32_t i = tid; i < NQ; i += SC_THREADS) {
    uint64_t t0 = 0;
    if (p.thr_mode && i < p.nq_valid) {
      // fixed cut-off: the largest key BELOW every key whose score is thr_cos * |q| or more
      const float t = p.thr_cos * p.qnorm[p.qmap ? p.qmap[i] : i];
      t0 = t == t ? ((uint64_t)ord_from_float(t) << 32) - 1ull : ~0ull;  // NaN norm: nominate nothing
    }
    tau_s[i] = t0;
    cnt_s[i] = 0;
    cur_s[i] = 0;
  }
  if (tid < 4) flag_s[tid] = 0;
  __syncthreads();

  const uint32_t my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

  if (warp == SC_CW) {
    // ---------------- producer ----------------
    if (lane == 0) {
      // Each consumer group owns its own sub-ring of S/2 stages, so a stage's barriers are
      // only ever waited on by one group and phases cannot be skipped (parity would alias).
      // (The bf16 variant's tiles are shared by both groups: one ring of S stages.)
      const uint32_t SG = HALF ? S : S / SC_GROUPS;
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t t = blockIdx.x + i * gridDim.x;
        const uint32_t r0 = t * TR;
        const uint32_t rows_here = min(TR, p.st.n_rows - r0);
        const uint32_t g = HALF ? 0u : i % SC_GROUPS, jbase = (HALF ? i : i / SC_GROUPS) * p.n_slices;
        for (uint32_t sl = 0; sl < p.n_slices; ++sl) {
          const uint32_t j = jbase + sl;
          const uint32_t stage = g * SG + (j % SG);
          if (j >= SG) mbar_wait(empty0 + 8 * stage, ((j / SG) - 1) & 1);
          const uint32_t dst = smem_u32(tiles + (size_t)stage * tile_bytes);
          const uint32_t k0 = sl * p.kslice;
          const uint32_t klen = min(p.kslice, ld - k0);
          const uint32_t bar = full0 + 8 * stage;
          const bool last = sl + 1 == p.n_slices && !HALF;  // the shadow's rows are normalised already
          const uint32_t rn_bytes = last ? ((rows_here + 3) & ~3u) * 4 : 0;
          const unsigned char* src =
              HALF ? reinterpret_cast<const unsigned char*>(p.st.E16) : reinterpret_cast<const unsigned char*>(p.st.E);
          if (p.n_slices == 1) {
            const uint32_t bytes = rows_here * ld * ES;
            mbar_arrive_expect_tx(bar, bytes + rn_bytes);
            bulk_g2s(dst, src + (size_t)r0 * ld * ES, bytes, bar);
          } else {
            mbar_arrive_expect_tx(bar, rows_here * klen * ES + rn_bytes);
            for (uint32_t r = 0; r < rows_here; ++r)
              bulk_g2s(dst + r * p.kslice * ES, src + ((size_t)(r0 + r) * ld + k0) * ES, klen * ES, bar);
          }
          if (last) bulk_g2s(smem_u32(rn_s + stage * SC_TILE_ROWS), p.st.rnorm + r0, rn_bytes, bar);
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const uint32_t ctid = tid;             // 0..SC_CT-1
  const uint32_t grp = warp / SC_GW;     // which alternate tiles
  const uint32_t gwarp = warp % SC_GW;   // row block inside the tile
  constexpr int NV = SC_R * NQ;
  const uint32_t idx = transpose_index<NV>(lane);
  const bool leader = (lane & ((32u / NV) - 1u)) == 0;  // NV <= 32
  const uint32_t my_r = idx / NQ, my_b = idx % NQ;
  const float4* q4 = reinterpret_cast<const float4*>(q_s);
  const uint32_t kslice4 = p.kslice >> 2;
  const uint32_t C = p.C, KP = p.KP;

  auto csync = [] { named_bar_sync(1, SC_CT); };
  auto gq = [&](uint32_t b) { return (p.qmap && b < p.nq_valid) ? p.qmap[b] : b; };  // state index of pass-local b

  // Keep the KP best of list[b] (rank by counting; keys are distinct), sorted, in the
  // other half of the double buffer; publish / adopt the global cut-off.
  auto compact = [&](uint32_t b) {
    const uint32_t n = min(cnt_s[b], C);
    const uint32_t cur = cur_s[b];
    const uint64_t* src = list_s + ((size_t)b * 2 + cur) * C;
    uint64_t* dst = list_s + ((size_t)b * 2 + (cur ^ 1)) * C;
    csync();
    for (uint32_t i = ctid; i < n; i += SC_CT) {
      const uint64_t key = src[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < n; ++j) rank += src[j] > key;
      if (rank < KP) dst[rank] = key;
    }
    csync();
    if (ctid == 0) {
      cnt_s[b] = min(n, KP);
      cur_s[b] = cur ^ 1;
      if (n >= KP) {
        uint64_t t = dst[KP - 1];
        const uint64_t g = atomicMax(reinterpret_cast<unsigned long long*>(p.gtau + gq(b)), (unsigned long long)t);
        if (g > t) t = g;
        if (t > tau_s[b]) tau_s[b] = t;
      }
    }
    csync();
  };

  // threshold mode: move list[b] to the query's merged list and start over
  auto flush = [&](uint32_t b) {
    csync();
    const uint32_t raw = cnt_s[b], n = min(raw, C);
    // a list that overran between two checkpoints lost keys: poison the count so that the
    // rescoring kernel reports an overflow (cannot happen while C >= 2 * SC_QUAD * SC_TILE_ROWS)
    if (ctid == 0) flag_s[2] = n ? atomicAdd(p.cnt + gq(b), n + (raw > C ? 0x40000000u : 0u)) : 0;
    csync();
    const uint32_t base = flag_s[2];
    const uint64_t* src = list_s + ((size_t)b * 2 + cur_s[b]) * C;
    uint64_t* out = p.keys + (size_t)b * p.cap;
    for (uint32_t j = ctid; j < n; j += SC_CT)
      if (base + j < p.cap) out[base + j] = src[j];  // excess is dropped; cnt > cap tells the rescoring kernel
    csync();
    if (ctid == 0) cnt_s[b] = 0;
    csync();
  };

  const uint32_t n_quads = my_tiles / QUAD;  // full quads: both groups see the same count
  uint32_t ck = 0;                           // bf16 variant: checkpoints passed
  // this warp's rows inside a tile, and its position in the ring (kept incrementally: no division per slice)
  const uint32_t rblock = (HALF ? grp * SC_GW + gwarp : gwarp) * SC_R;
  const uint32_t SGc = HALF ? S : S / SC_GROUPS, stage0 = HALF ? 0u : grp * SGc;
  uint32_t slot = 0, phase = 0;

  for (uint32_t i = HALF ? 0u : grp; i < my_tiles; i += HALF ? 1u : (uint32_t)SC_GROUPS) {
    const uint32_t t = blockIdx.x + i * gridDim.x;
    float acc[SC_R][NQ];
#pragma unroll
    for (int r = 0; r < SC_R; ++r)
#pragma unroll
      for (int b = 0; b < NQ; ++b) acc[r][b] = 0.0f;

    float my_rn = 0.0f;
    for (uint32_t sl = 0; sl < p.n_slices; ++sl) {
      const uint32_t stage = stage0 + slot;  // this group's sub-ring (see the producer)
      mbar_wait(full0 + 8 * stage, phase);
      if (++slot == SGc) {
        slot = 0;
        phase ^= 1u;
      }
      // four elements per lane and load: a float4 of an fp32 row, a uint2 (4 x bf16) of a shadow row
      const unsigned char* tile_b = tiles + (size_t)stage * tile_bytes + (size_t)rblock * p.kslice * ES;
      auto load4 = [&](uint32_t r, uint32_t c4) -> float4 {
        if (HALF) {
          const uint2 u = reinterpret_cast<const uint2*>(tile_b)[r * kslice4 + c4];
          return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                             __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
        }
        return reinterpret_cast<const float4*>(tile_b)[r * kslice4 + c4];
      };
      const uint32_t k0_4 = (sl * p.kslice) >> 2;
      if (KC > 0) {
#pragma unroll
        for (int c = 0; c < (KC > 0 ? KC : 1); ++c) {
          float4 e[SC_R];
#pragma unroll
          for (int r = 0; r < SC_R; ++r) e[r] = load4(r, c * 32 + lane);
#pragma unroll
          for (int b = 0; b < NQ; ++b) {
            const float4 qv = q4[b * ld4 + k0_4 + c * 32 + lane];
#pragma unroll
            for (int r = 0; r < SC_R; ++r) {
              acc[r][b] = fmaf(e[r].x, qv.x, acc[r][b]);
              acc[r][b] = fmaf(e[r].y, qv.y, acc[r][b]);
              acc[r][b] = fmaf(e[r].z, qv.z, acc[r][b]);
              acc[r][b] = fmaf(e[r].w, qv.w, acc[r][b]);
            }
          }
        }
      } else {
        const uint32_t klen4 = min(kslice4, ld4 - k0_4);
        for (uint32_t c = lane; c < klen4; c += 32) {
          float4 e[SC_R];
#pragma unroll
          for (int r = 0; r < SC_R; ++r) e[r] = load4(r, c);
#pragma unroll
          for (int b = 0; b < NQ; ++b) {
            const float4 qv = q4[b * ld4 + k0_4 + c];
#pragma unroll
            for (int r = 0; r < SC_R; ++r) {
              acc[r][b] = fmaf(e[r].x, qv.x, acc[r][b]);
              acc[r][b] = fmaf(e[r].y, qv.y, acc[r][b]);
              acc[r][b] = fmaf(e[r].z, qv.z, acc[r][b]);
              acc[r][b] = fmaf(e[r].w, qv.w, acc[r][b]);
            }
          }
        }
      }
      if (sl + 1 == p.n_slices) my_rn = HALF ? 1.0f : rn_s[stage * SC_TILE_ROWS + gwarp * SC_R + my_r];
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);
    }

    float v[NV];
#pragma unroll
    for (int r = 0; r < SC_R; ++r)
#pragma unroll
      for (int b = 0; b < NQ; ++b) v[r * NQ + b] = acc[r][b];
    const float total = warp_transpose_reduce<NV>(v, lane);

    const uint32_t row = t * TR + rblock + my_r;
    if (leader && row < p.st.n_rows && my_b < p.nq_valid) {
      // cosine up to the query's own (positive) norm, which cannot change the order
      // within a query; cx_select.cu applies it when it compares against eps
      const float approx = total * my_rn;
      if (approx == approx) {
        const uint64_t key = make_key(ord_from_float(approx), row);
        if (key > *((volatile uint64_t*)(tau_s + my_b))) {
          if (row_passes(p.flt, p.st.meta, p.st.agent, row)) {
            const uint32_t pos = atomicAdd(cnt_s + my_b, 1u);
            const uint32_t cur = *((volatile uint32_t*)(cur_s + my_b));
            if (pos < C) list_s[((size_t)my_b * 2 + cur) * C + pos] = key;
          }
        }
      }
    }

    // List check.  fp32 variant: after the last tile of every full quad (tile i%4 == 2 for group 0, 3 for group 1).
    // bf16 variant: every second tile at first, then every QUAD-th (every warp walks every tile, so all agree).
    bool check;
    if (HALF) check = (i < SC_EARLY_TILES ? (i & 1u) == 1u : ((i - SC_EARLY_TILES) % QUAD) == QUAD - 1) && i + 1 < my_tiles;
    else check = (i % QUAD) == (QUAD - SC_GROUPS + grp) && (i / QUAD) < n_quads;
    if (check) {
      if (HALF) {
        // One barrier per checkpoint: which lists to compact was decided (by thread 0) right after the PREVIOUS
        // checkpoint, from counts that can only have been too high; a list is flagged while two more quads of
        // appends might not fit, so it can never overrun before the flag takes effect (C >= 2 KP + 2 quads).
        csync();
        const uint32_t m = flag_s[(ck & 1u) ? 3 : 0];
        if (ctid < NQ) {
          const uint64_t g = *((volatile uint64_t*)(p.gtau + gq(ctid)));
          if (g > tau_s[ctid]) tau_s[ctid] = g;
        }
        for (uint32_t b = 0; b < NQ; ++b)
          if (m & (1u << b)) compact(b);
        if (ctid == 0) {
          uint32_t nm = 0;
          for (uint32_t b = 0; b < NQ; ++b) {
            const uint32_t c = *((volatile uint32_t*)(cnt_s + b));
            if (c >= 2 * KP || c + 2 * QUAD * TR > C) nm |= 1u << b;
          }
          flag_s[((ck + 1) & 1u) ? 3 : 0] = nm;
        }
        ++ck;
      } else {
        csync();
        if (ctid < NQ && !p.thr_mode) {
          const uint64_t g = *((volatile uint64_t*)(p.gtau + gq(ctid)));
          if (g > tau_s[ctid]) tau_s[ctid] = g;
        }
        if (ctid == 0) {
          uint32_t m = 0;
          for (uint32_t b = 0; b < NQ; ++b)
            if ((!p.thr_mode && cnt_s[b] >= 2 * KP) || cnt_s[b] + SC_QUAD * SC_TILE_ROWS > C) m |= 1u << b;
          *flag_s = m;
        }
        csync();
        const uint32_t m = *flag_s;
        for (uint32_t b = 0; b < NQ; ++b)
          if (m & (1u << b)) {
            if (p.thr_mode) flush(b);
            else compact(b);
          }
      }
    }
  }

  if (p.thr_mode) {
    for (uint32_t b = 0; b < p.nq_valid; ++b) flush(b);
    return;
  }

  // final: rank every list and append the part that can still matter -- keys at or
  // above the global cut-off -- to the query's merged list.  gtau ends up as the
  // maximum of every group's cut-off, i.e. an upper bound of everything dropped.
  csync();
  for (uint32_t b = 0; b < p.nq_valid; ++b) {
    compact(b);
    const uint64_t* Lb = list_s + ((size_t)b * 2 + cur_s[b]) * C;
    const uint32_t n_keep = cnt_s[b];
    if (ctid == 0) {
      const uint64_t g = *((volatile uint64_t*)(p.gtau + gq(b)));
      uint32_t m = 0;
      while (m < n_keep && Lb[m] >= g) ++m;  // sorted descending: a prefix survives
      flag_s[1] = m;
      flag_s[2] = m ? atomicAdd(p.cnt + gq(b), m) : 0;
    }
    csync();
    const uint32_t m = flag_s[1], base = flag_s[2];
    uint64_t* out = p.keys + (size_t)b * p.cap + base;
    for (uint32_t j = ctid; j < m; j += SC_CT)
      if (base + j < p.cap) out[j] = Lb[j];  // excess is dropped; cnt > cap tells the select kernel
    csync();
  }
}

uint32_t stream_scan_groups(uint32_t n_rows, int sm_count) {
  uint32_t n_tiles = (n_rows + SC_TILE_ROWS - 1) / SC_TILE_ROWS;
  return n_tiles < (uint32_t)sm_count ? n_tiles : (uint32_t)sm_count;
}

static uint32_t pow2_at_least(uint32_t x) {
  uint32_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

// floats of a row per stage: the whole row if it fits 48 KB stages, else the largest
// multiple of 128 that divides the row (unrolled kernel variant), else 384
static uint32_t stream_pick_kslice(uint32_t ld) {
  if (ld <= (uint32_t)SC_KSLICE) return ld;
  for (uint32_t ks = SC_KSLICE; ks >= 128; ks -= 128)
    if (ld % ks == 0) return ks;
  return SC_KSLICE;
}

// entries of a per-query list in shared memory: the kept keys twice over plus what can arrive between two
// checks (one quad = 128 rows; the bf16 variant decides one check ahead, so two)
static uint32_t stream_list_cap(uint32_t KP, bool half, uint32_t quad = 2) {
  if (half) return (2 * KP + 2 * quad * sc_tile_rows(true) + 63) & ~63u;
  return pow2_at_least(2 * KP + SC_QUAD * SC_TILE_ROWS);
}

static uint32_t stream_pick_stages(uint32_t ld, uint32_t nq, uint32_t KP, uint32_t esize, size_t* total,
                                   uint32_t quad = 2) {
  uint32_t kslice = stream_pick_kslice(ld);
  const bool half = esize == 2;
  uint32_t C = stream_list_cap(KP, half, quad);
  // fp32 variant: even (each consumer group owns s/2 stages); bf16 variant: one ring shared by both groups
  for (uint32_t s = SC_MAX_STAGES; s >= 2; s -= half ? 1 : SC_GROUPS) {
    StreamLayout L = stream_layout(ld, nq, C, s, kslice, esize);
    if (L.total <= SC_SMEM_LIMIT) {
      *total = L.total;
      return s;
    }
  }
  *total = 0;
  return 0;
}

size_t stream_scan_smem(uint32_t ld, uint32_t nq_pass, uint32_t KP) {
  size_t total;
  uint32_t s = stream_pick_stages(ld, nq_pass, KP, 4, &total);
  return s ? total : 0;
}

// the bf16-shadow variant: ld16 elements per row, at most four queries per pass
size_t stream_scan_half_smem(uint32_t ld16, uint32_t nq_pass, uint32_t KP) {
  size_t total;
  if (nq_pass > 4) return 0;
  uint32_t s = stream_pick_stages(ld16, nq_pass, KP, 2, &total);
  return s ? total : 0;
}

template <int NQ, int KC, bool HALF>
static cudaError_t launch_one(const StreamParams& p, size_t smem, cudaStream_t s) {
  cudaError_t e = raise_dynamic_smem<stream_scan_kernel<NQ, KC, HALF>>(smem);
  if (e != cudaSuccess) return e;
  stream_scan_kernel<NQ, KC, HALF><<<p.G, SC_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

template <int NQ, bool HALF>
static cudaError_t launch_nq(const StreamParams& p, size_t smem, cudaStream_t s) {
  // unrolled variants for slices that are exactly 1, 2 or 3 chunks of 128 elements
  const uint32_t ld = HALF ? p.st.ld16 : p.st.ld;
  const bool even = (ld % p.kslice) == 0 && (p.kslice % 128) == 0;
  const uint32_t kc = even ? p.kslice / 128 : 0;
  switch (kc) {
    case 1: return launch_one<NQ, 1, HALF>(p, smem, s);
    case 2: return launch_one<NQ, 2, HALF>(p, smem, s);
    case 3: return launch_one<NQ, 3, HALF>(p, smem, s);
    default: return launch_one<NQ, 0, HALF>(p, smem, s);
  }
}

cudaError_t launch_stream_scan(const StoreView& st, const QueryView& qv, uint32_t q0, uint32_t nq_pass,
                               const DevFilter& flt, const CandView& cv, int sm_count, cudaStream_t s,
                               const uint32_t* qmap, const float* thr_cos, bool half) {
  if (!st.n_rows || !nq_pass) return cudaSuccess;
  if (qmap && q0 != 0) return cudaErrorInvalidValue;
  if (nq_pass > (half ? 4u : 8u)) return cudaErrorInvalidValue;
  if (half && (!st.E16 || thr_cos)) return cudaErrorInvalidValue;
  const uint32_t nq_t = nq_pass <= 1 ? 1 : nq_pass <= 2 ? 2 : nq_pass <= 4 ? 4 : 8;
  const uint32_t ld = half ? st.ld16 : st.ld;
  StreamParams p;
  p.st = st;
  p.Q = qv.Q + (size_t)q0 * qv.ldq;
  p.ldq = qv.ldq;
  p.nq_valid = nq_pass;
  p.flt = flt;
  p.G = cv.G;
  p.KP = cv.KP;
  p.cap = cv.cap;
  p.qmap = qmap;
  p.thr_mode = thr_cos ? 1u : 0u;
  p.thr_cos = thr_cos ? *thr_cos : 0.0f;
  p.qnorm = qmap ? qv.qnorm : qv.qnorm + q0;
  p.keys = cv.keys + (size_t)(q0 - cv.q_base) * cv.cap;
  p.cnt = cv.cnt + q0;
  p.gtau = cv.gtau + q0;
  p.C = stream_list_cap(cv.KP, half);
  p.n_tiles = (st.n_rows + sc_tile_rows(half) - 1) / sc_tile_rows(half);
  p.kslice = stream_pick_kslice(ld);
  p.n_slices = (ld + p.kslice - 1) / p.kslice;
  size_t smem;
  p.stages = stream_pick_stages(ld, nq_t, cv.KP, half ? 2 : 4, &smem);
  if (!p.stages) return cudaErrorInvalidConfiguration;
  p.quad = 2;
  if (half) {
    // rarer list checks (each is a barrier over all consumer warps) while the longer lists cost no ring stage
    for (uint32_t quad = 8; quad > 2; quad >>= 1) {
      size_t t;
      if (stream_pick_stages(ld, nq_t, cv.KP, 2, &t, quad) == p.stages) {
        p.quad = quad;
        smem = t;
        break;
      }
    }
    p.C = stream_list_cap(cv.KP, true, p.quad);
  }
  if (p.G != stream_scan_groups(st.n_rows, sm_count)) return cudaErrorInvalidValue;
  if (half) {
    switch (nq_t) {
      case 1: return launch_nq<1, true>(p, smem, s);
      case 2: return launch_nq<2, true>(p, smem, s);
      default: return launch_nq<4, true>(p, smem, s);
    }
  }
  switch (nq_t) {
    case 1: return launch_nq<1, false>(p, smem, s);
    case 2: return launch_nq<2, false>(p, smem, s);
    case 4: return launch_nq<4, false>(p, smem, s);
    default: return launch_nq<8, false>(p, smem, s);
  }
}

}  // namespace cx
