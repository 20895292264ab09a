// cx_stream.cu -- K1: small-batch streaming fp32 scan (HBM-bound).
//
// One persistent CTA per SM.  A producer thread streams 32-row tiles of the
// row-major embedding matrix into a shared-memory ring with 1-D bulk async
// copies (cp.async.bulk -> UBLKCP, completion on mbarriers); eight consumer
// warps take four rows each, lanes split the dimension (conflict-free LDS.128),
// accumulate NQ query dot products per row with FFMA, and finish with a
// transposing warp reduction so every lane ends up owning one (row, query)
// score.  Scores only nominate candidates: each CTA keeps a per-query list in
// shared memory with a running cut-off (the KP-th best key seen so far), and the
// lists are merged, rescored with reference arithmetic and verified by
// cx_select.cu.  The matrix is read exactly once per pass:
//   algorithmic bytes per pass = n_rows * ld * 4   
//
// Call shapes served (reference): search() at B=1 (api.rs:117-125,
// http/routes.rs:906-907, grpc/service.rs:673-675, gate/mod.rs:332) and small
// search_batch() groups (vector/index.rs:390-410).
#include "cx_kernels.h"

namespace cx {

constexpr int SC_W = 8;                       // consumer warps
constexpr int SC_THREADS = 32 * (SC_W + 1);   // + 1 producer warp
constexpr int SC_R = 4;                       // rows per warp per tile
constexpr int SC_TILE_ROWS = SC_W * SC_R;     // 32
constexpr int SC_CHECK = 4;                   // tiles between list checks
constexpr int SC_KSLICE = 384;                // floats of a row per stage (<= 48 KB stages)
constexpr int SC_MAX_STAGES = 4;
constexpr size_t SC_SMEM_LIMIT = 227 * 1024;

struct StreamParams {
  StoreView st;
  const float* Q;    // first query of this pass
  const float* rqn;  // its reciprocal norm
  uint32_t ldq, nq_valid;
  DevFilter flt;
  uint64_t* keys;    // &cand.keys[q0][0][0]
  uint64_t* bound;   // &cand.bound[q0][0]
  uint32_t G, KP, C, n_tiles, stages, kslice, n_slices;
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

__host__ __device__ inline size_t stream_layout(uint32_t ld, uint32_t nq, uint32_t C, uint32_t stages,
                                                uint32_t kslice, size_t* off_q, size_t* off_list,
                                                size_t* off_tau, size_t* off_cnt, size_t* off_bars) {
  size_t o = (size_t)stages * SC_TILE_ROWS * kslice * 4;
  *off_q = o;
  o += (size_t)nq * ld * 4;
  o = (o + 15) & ~(size_t)15;
  *off_list = o;
  o += (size_t)nq * C * 8;
  *off_tau = o;
  o += (size_t)nq * 8;
  *off_cnt = o;
  o += (size_t)(nq + 4) * 4;  // cnt[nq] + flag
  o = (o + 7) & ~(size_t)7;
  *off_bars = o;
  o += (size_t)2 * SC_MAX_STAGES * 8;
  return o;
}

template <int NQ>
__global__ void __launch_bounds__(SC_THREADS, 1) stream_scan_kernel(const StreamParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  size_t off_q, off_list, off_tau, off_cnt, off_bars;
  stream_layout(p.st.ld, NQ, p.C, p.stages, p.kslice, &off_q, &off_list, &off_tau, &off_cnt, &off_bars);
  float* tiles = reinterpret_cast<float*>(smem_raw);
  float* q_s = reinterpret_cast<float*>(smem_raw + off_q);
  uint64_t* list_s = reinterpret_cast<uint64_t*>(smem_raw + off_list);
  uint64_t* tau_s = reinterpret_cast<uint64_t*>(smem_raw + off_tau);
  uint32_t* cnt_s = reinterpret_cast<uint32_t*>(smem_raw + off_cnt);
  uint32_t* flag_s = cnt_s + NQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + off_bars);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ld = p.st.ld, ld4 = ld >> 2, S = p.stages;
  const uint32_t tile_floats = SC_TILE_ROWS * p.kslice;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + SC_MAX_STAGES);

  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, SC_W);
    }
    fence_mbar_init();
  }
  // queries -> smem (stride ld), lists cleared
  for (uint32_t i = tid; i < NQ * ld; i += SC_THREADS) {
    uint32_t b = i / ld, d = i % ld;
    q_s[i] = (b < p.nq_valid) ? p.Q[(size_t)b * p.ldq + d] : 0.0f;
  }
  for (uint32_t i = tid; i < NQ; i += SC_THREADS) {
    tau_s[i] = 0;
    cnt_s[i] = 0;
  }
  __syncthreads();

  const uint32_t my_tiles = (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == SC_W) {
    // ---------------- producer ----------------
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t t = blockIdx.x + i * gridDim.x;
        const uint32_t r0 = t * SC_TILE_ROWS;
        const uint32_t rows_here = min((uint32_t)SC_TILE_ROWS, p.st.n_rows - r0);
        for (uint32_t sl = 0; sl < p.n_slices; ++sl, ++it) {
          const uint32_t stage = it % S;
          if (it >= S) mbar_wait(empty0 + 8 * stage, ((it / S) - 1) & 1);
          const uint32_t dst = smem_u32(tiles + (size_t)stage * tile_floats);
          const uint32_t k0 = sl * p.kslice;
          const uint32_t klen = min(p.kslice, ld - k0);
          const uint32_t bar = full0 + 8 * stage;
          if (p.n_slices == 1) {
            const uint32_t bytes = rows_here * ld * 4;
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(dst, p.st.E + (size_t)r0 * ld, bytes, bar);
          } else {
            mbar_arrive_expect_tx(bar, rows_here * klen * 4);
            for (uint32_t r = 0; r < rows_here; ++r)
              bulk_g2s(dst + r * p.kslice * 4, p.st.E + (size_t)(r0 + r) * ld + k0, klen * 4, bar);
          }
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const uint32_t ctid = tid;  // 0..255
  constexpr int NV = SC_R * NQ;
  const uint32_t idx = transpose_index<NV>(lane);
  const bool leader = (lane & ((32u / NV) - 1u)) == 0;  // NV <= 32
  const uint32_t my_r = idx / NQ, my_b = idx % NQ;
  const float my_rqn = (my_b < p.nq_valid) ? __ldg(p.rqn + my_b) : 0.0f;
  const float4* q4 = reinterpret_cast<const float4*>(q_s);
  const uint32_t kslice4 = p.kslice >> 2;

  auto csync = [] { named_bar_sync(1, SC_W * 32); };

  auto compact = [&](uint32_t b, bool final_pass) {
    // all consumer threads; list[b] sorted descending, truncated to KP
    uint64_t* L = list_s + (size_t)b * p.C;
    uint32_t n = min(cnt_s[b], p.C);
    csync();
    for (uint32_t i = ctid; i < p.C; i += SC_W * 32)
      if (i >= n) L[i] = 0;
    csync();
    bitonic_sort_desc(L, p.C, ctid, SC_W * 32, csync);
    if (ctid == 0) {
      if (n > p.KP) {
        tau_s[b] = L[p.KP - 1];
        cnt_s[b] = p.KP;
      }
    }
    (void)final_pass;
    csync();
  };

  uint32_t it = 0;
  for (uint32_t i = 0; i < my_tiles; ++i) {
    const uint32_t t = blockIdx.x + i * gridDim.x;
    float acc[SC_R][NQ];
#pragma unroll
    for (int r = 0; r < SC_R; ++r)
#pragma unroll
      for (int b = 0; b < NQ; ++b) acc[r][b] = 0.0f;

    for (uint32_t sl = 0; sl < p.n_slices; ++sl, ++it) {
      const uint32_t stage = it % S;
      mbar_wait(full0 + 8 * stage, (it / S) & 1);
      const float4* tile4 = reinterpret_cast<const float4*>(tiles + (size_t)stage * tile_floats);
      const uint32_t k0_4 = (sl * p.kslice) >> 2;
      const uint32_t klen4 = min(kslice4, ld4 - k0_4);
      for (uint32_t c = lane; c < klen4; c += 32) {
        float4 e[SC_R];
#pragma unroll
        for (int r = 0; r < SC_R; ++r) e[r] = tile4[(warp * SC_R + r) * kslice4 + c];
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
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);
    }

    float v[NV];
#pragma unroll
    for (int r = 0; r < SC_R; ++r)
#pragma unroll
      for (int b = 0; b < NQ; ++b) v[r * NQ + b] = acc[r][b];
    const float total = warp_transpose_reduce<NV>(v, lane);

    const uint32_t row = t * SC_TILE_ROWS + warp * SC_R + my_r;
    if (leader && row < p.st.n_rows && my_b < p.nq_valid) {
      const float approx = total * __ldg(p.st.rnorm + row) * my_rqn;
      if (approx == approx) {
        const uint64_t key = make_key(ord_from_float(approx), row);
        if (key > *((volatile uint64_t*)(tau_s + my_b))) {
          if (row_passes(p.flt, p.st.meta, p.st.agent, row)) {
            uint32_t pos = atomicAdd(cnt_s + my_b, 1u);
            if (pos < p.C) list_s[(size_t)my_b * p.C + pos] = key;
          }
        }
      }
    }

    if ((i + 1) % SC_CHECK == 0 && i + 1 < my_tiles) {
      csync();
      if (ctid == 0) {
        uint32_t m = 0;
        for (uint32_t b = 0; b < NQ; ++b)
          if (cnt_s[b] + SC_CHECK * SC_TILE_ROWS > p.C) m |= 1u << b;
        *flag_s = m;
      }
      csync();
      const uint32_t m = *flag_s;
      for (uint32_t b = 0; b < NQ; ++b)
        if (m & (1u << b)) compact(b, false);
    }
  }

  // final: sort every list, emit the head
  csync();
  for (uint32_t b = 0; b < p.nq_valid; ++b) {
    const uint32_t n_before = min(cnt_s[b], p.C);
    const uint64_t tau_before = tau_s[b];
    compact(b, true);
    const uint64_t* L = list_s + (size_t)b * p.C;
    uint64_t* out = p.keys + ((size_t)b * p.G + blockIdx.x) * p.KP;
    const uint32_t n_keep = min(n_before, p.KP);
    for (uint32_t j = ctid; j < p.KP; j += SC_W * 32) out[j] = j < n_keep ? L[j] : 0;
    if (ctid == 0) p.bound[(size_t)b * p.G + blockIdx.x] = (n_before > p.KP) ? L[p.KP - 1] : tau_before;
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

static uint32_t stream_list_cap(uint32_t KP) { return pow2_at_least(KP + 2 * SC_CHECK * SC_TILE_ROWS); }

static uint32_t stream_pick_stages(uint32_t ld, uint32_t nq, uint32_t KP, size_t* total) {
  uint32_t kslice = ld < (uint32_t)SC_KSLICE ? ld : SC_KSLICE;
  uint32_t C = stream_list_cap(KP);
  for (uint32_t s = SC_MAX_STAGES; s >= 2; --s) {
    size_t a, b, c, d, e;
    size_t bytes = stream_layout(ld, nq, C, s, kslice, &a, &b, &c, &d, &e);
    if (bytes <= SC_SMEM_LIMIT) {
      *total = bytes;
      return s;
    }
  }
  *total = 0;
  return 0;
}

size_t stream_scan_smem(uint32_t ld, uint32_t nq_pass, uint32_t KP) {
  size_t total;
  uint32_t s = stream_pick_stages(ld, nq_pass, KP, &total);
  return s ? total : 0;
}

template <int NQ>
static cudaError_t launch_nq(const StreamParams& p, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(stream_scan_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  stream_scan_kernel<NQ><<<p.G, SC_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_stream_scan(const StoreView& st, const QueryView& qv, uint32_t q0, uint32_t nq_pass,
                               const DevFilter& flt, const CandView& cv, int sm_count, cudaStream_t s) {
  if (!st.n_rows || !nq_pass) return cudaSuccess;
  uint32_t nq_t = nq_pass <= 1 ? 1 : nq_pass <= 2 ? 2 : nq_pass <= 4 ? 4 : 8;
  if (nq_pass > 8) return cudaErrorInvalidValue;
  StreamParams p;
  p.st = st;
  p.Q = qv.Q + (size_t)q0 * qv.ldq;
  p.rqn = qv.rqnorm + q0;
  p.ldq = qv.ldq;
  p.nq_valid = nq_pass;
  p.flt = flt;
  p.G = cv.G;
  p.KP = cv.KP;
  p.keys = cv.keys + (size_t)q0 * cv.G * cv.KP;
  p.bound = cv.bound + (size_t)q0 * cv.G;
  p.C = stream_list_cap(cv.KP);
  p.n_tiles = (st.n_rows + SC_TILE_ROWS - 1) / SC_TILE_ROWS;
  p.kslice = st.ld < (uint32_t)SC_KSLICE ? st.ld : SC_KSLICE;
  p.n_slices = (st.ld + p.kslice - 1) / p.kslice;
  size_t smem;
  p.stages = stream_pick_stages(st.ld, nq_t, cv.KP, &smem);
  if (!p.stages) return cudaErrorInvalidConfiguration;
  if (p.G != stream_scan_groups(st.n_rows, sm_count)) return cudaErrorInvalidValue;
  switch (nq_t) {
    case 1: return launch_nq<1>(p, smem, s);
    case 2: return launch_nq<2>(p, smem, s);
    case 4: return launch_nq<4>(p, smem, s);
    default: return launch_nq<8>(p, smem, s);
  }
}

}  // namespace cx
