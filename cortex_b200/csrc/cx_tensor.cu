// cx_tensor.cu -- K2: batched Q.E^T on the 5th-gen tensor cores (tcgen05 + TMEM),
// operands fed by TMA, with the candidate selection fused into the epilogue so the
// score matrix is never written.
//
// Shapes (reference call sites): search_batch (vector/index.rs:390-410), the
// auto-linker's per-node search loop (linker/auto_linker.rs:215-222) and the dedup
// self-join (linker/dedup.rs:72-88) all become  scores[B, N] = Qn[B, D] . En[N, D]^T
// on the NORMALISED bf16 copies (so a score is an approximate cosine).
//
// One CTA per SM, 12 warps:
//   warp 0  TMA producer: the CTA's 128-query tile once (resident in smem for the
//           whole pass), then [256 rows x 64] bf16 boxes of E through a ring of stages
//   warp 1  one thread issues tcgen05.mma (M=128 queries x N=256 rows x K=16),
//           accumulators in TMEM, double buffered (2 x 256 columns)
//   warp 2  TMEM allocation
//   warps 4-11 epilogue: lane == query, two warps per TMEM lane quarter (each takes half
//           of the 256 columns), tcgen05.ld 32 columns at a time.
//
// The pass runs in two modes over the same pipeline:
//   DUMP  scores of a strided sample of row tiles are written out; tau_select_kernel
//         turns them into a per-query cut-off tau0 = KP-th best sampled score (a valid
//         lower bound of the KP-th best over the whole corpus)
//   SCAN  every row; the common path of the epilogue is a running max against the
//         query's cut-off.  Scores at or above it are appended to the thread's private
//         list (global scratch, L2 resident).  A list that fills up is reduced to its
//         best KP by the warp cooperatively (bitonic sort in registers), which raises
//         that thread's cut-off and publishes it.
// CTAs are arranged (query tile) x (row split); CTAs that differ only in the query
// tile walk the same rows at the same time, so E streams from HBM once and is served
// to the others from L2.
//
// Algorithmic work: 2*D flops per scored pair; DESIGN.md §4.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>

#include "cx_kernels.h"

namespace cx {

constexpr uint32_t TC_BM = 128;       // queries per CTA (UMMA M, TMEM lanes)
constexpr uint32_t TC_BN = 256;       // corpus rows per tile (UMMA N, TMEM columns)
constexpr uint32_t TC_BK = 64;        // bf16 per K chunk = one 128 B swizzle atom
constexpr uint32_t TC_UK = 16;        // K of one tcgen05.mma for 16-bit inputs
constexpr uint32_t TC_CTRL_WARPS = 4;     // TMA producer, MMA issuer, TMEM allocator, (idle)
constexpr uint32_t TC_LIST_POOL = 256 * 512;  // private-list entries per CTA, split among its epilogue threads
constexpr uint32_t TC_QCHUNK_BYTES = TC_BM * TC_BK * 2;  // 16 KB
constexpr uint32_t TC_MAX_STAGES = 8;
// bytes of one ring stage per CTA: the whole [256 x 64] bf16 box, or -- for a CTA pair, where
// each CTA stages half of the rows and the pair's MMA reads both halves -- [128 x 64]
__host__ __device__ constexpr uint32_t tc_stage_bytes(bool pair) { return (pair ? TC_BN / 2 : TC_BN) * TC_BK * 2; }
constexpr uint32_t TC_SMEM_LIMIT = 227 * 1024;
constexpr uint32_t TC_TMEM_COLS = 512;
// EW epilogue warps (8 or 16): four per TMEM lane quarter share the 256 columns of a tile;
// private list entries per epilogue thread and per lane in a cooperative sort
__host__ __device__ constexpr uint32_t tc_list_cap(uint32_t EW) { return TC_LIST_POOL / (EW * 32); }

enum { TC_MODE_SCAN = 0, TC_MODE_DUMP = 1 };

struct TensorParams {
  uint32_t n_rows, n_tiles, n_kc;
  uint32_t n_slots;     // tiles visited: a contiguous tile range in SCAN mode, the sample size in DUMP mode
  uint32_t tile0;       // SCAN mode: first tile of this launch's range
  uint32_t n_qt, n_es;
  uint32_t n_tail, n_rem;  // the last n_tail slots belong to the n_rem CTAs left over by the n_qt x n_es grid
  uint32_t nq_valid;
  uint32_t stages;
  uint32_t mode;
  uint32_t check_rows;  // 0: no filter and no removed rows -> skip the per-row metadata test
  uint32_t static_tau;  // threshold scans: the cut-off in gtau is fixed; a full private list is flushed to the
                        // query's merged list instead of being cut to its best KP
  uint32_t debug;       // measurement hook (results become wrong): 1 = epilogue only releases the
                        // accumulator, 2 = epilogue loads and masks but never handles a hit
  uint32_t KP;
  const uint32_t* meta;
  const uint32_t* agent;
  DevFilter flt;
  uint64_t* lists;  // [grid][TC_LIST_POOL] private candidate lists (scratch), one slice per epilogue thread
  uint64_t* keys;   // merged list per query [nq][cap]
  uint32_t* cnt;
  uint64_t* gtau;
  uint32_t cap;
  float* dump;      // DUMP mode: [nq][n_slots * 256] sampled scores (-inf where not eligible)
};

// ---- PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// CTA-pair form: the mbarrier may live in the peer CTA (a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address -> shared::cluster address of the same location in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// pair forms: executed by the same warp of both CTAs
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M=256 over the pair (128 queries per CTA) x N=256 (128 rows staged by each CTA); leader CTA only
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same offset in both CTAs of the pair when the MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128 B swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4, LBO (unused for swizzled K-major) = 1, SBO = 1024 B (8 rows of
// 128 B), version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((TC_BN >> 3) << 17) | ((TC_BM >> 4) << 24);
constexpr uint32_t TC_IDESC_PAIR = (1u << 4) | (1u << 7) | (1u << 10) | ((TC_BN >> 3) << 17) | (((2 * TC_BM) >> 4) << 24);

// v[j] for a run-time j without spilling v to local memory: 31 selects
__device__ __forceinline__ float pick32(const float (&v)[32], uint32_t j) {
  float a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1u) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2u) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4u) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8u) ? c[2 * i + 1] : c[2 * i];
  return (j & 16u) ? d[1] : d[0];
}

// ---- warp-cooperative reduction of one private list --------------------------------
// Sort (descending) the n <= 32 * NPL keys at L across the warp's registers, write the best KP
// back in order, return the KP-th key (0 if n < KP).  All lanes call it.
template <int NPL>
__device__ __noinline__ uint64_t warp_compact(uint64_t* L, uint32_t n, uint32_t KP, uint32_t lane) {
  uint64_t k[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const uint32_t e = i * 32 + lane;
    k[i] = e < n ? L[e] : 0ull;
  }
  constexpr uint32_t N = NPL * 32;
#pragma unroll
  for (uint32_t kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
    for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const uint32_t ji = j >> 5;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          const int pi = i ^ (int)ji;
          if (pi > i) {
            const uint32_t e = i * 32 + lane;
            const bool desc = (e & kk) == 0;
            const uint64_t a = k[i], b = k[pi];
            const bool sw = desc ? (a < b) : (a > b);
            k[i] = sw ? b : a;
            k[pi] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          const uint32_t e = i * 32 + lane;
          const uint64_t mine = k[i];
          const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, j);
          const bool lower = (lane & j) == 0;  // I hold the lower index of the pair
          const bool desc = (e & kk) == 0;     // descending block: lower index keeps the larger key
          const bool keep_max = (lower == desc);
          k[i] = keep_max ? (mine > other ? mine : other) : (mine < other ? mine : other);
        }
      }
    }
  }
  uint64_t kth = 0ull;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const uint32_t e = i * 32 + lane;
    if (e < KP && e < n) L[e] = k[i];
    const uint64_t cand = __shfl_sync(0xffffffffu, k[i], (KP - 1) & 31);
    if ((uint32_t)i == ((KP - 1) >> 5)) kth = cand;
  }
  __syncwarp();
  return n >= KP ? kth : 0ull;
}

// measurement hook: SM cycles and nanoseconds block 0 spent in its last launch (-> effective clock)
__device__ unsigned long long g_tensor_clock[2];

// PAIR = two CTAs of a cluster (one TPC) work as one tcgen05 cta_group::2 unit: each holds
// its own 128-query tile (the pair's M = 256) and stages HALF of every [256 x 64] box of E;
// the leader CTA issues the MMAs, which read both halves.  Per SM that halves the L2->SM
// traffic and the shared-memory reads of the E operand (the binding limits of the
// single-CTA form: 64 B/clk of TMA writes + 96 B/clk of MMA operand reads against
// 128 B/clk of shared-memory bandwidth).
// QRES = the CTA's query tile stays resident in shared memory for the whole pass (dimensions up to
// 640); otherwise (large embeddings, e.g. 1024-d) every ring stage carries the matching [128 x 64]
// chunk of the query tile next to the chunk of E, re-read from L2 for every row tile.
template <bool PAIR, int EW, bool QRES>
__global__ void __launch_bounds__((TC_CTRL_WARPS + EW) * 32, 1)
tensor_scan_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmE,
                   const TensorParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr uint32_t E_BYTES = tc_stage_bytes(PAIR);                           // E part of a stage
  constexpr uint32_t STAGE_BYTES = E_BYTES + (QRES ? 0u : TC_QCHUNK_BYTES);  // [E chunk][Q chunk]
  const uint32_t S = p.stages;
  unsigned char* sQ = smem;
  unsigned char* sE = smem + (QRES ? (size_t)p.n_kc * TC_QCHUNK_BYTES : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sE + (size_t)S * STAGE_BYTES);
  const uint32_t bar_q = smem_u32(bars);
  const uint32_t bar_full = smem_u32(bars + 1);
  const uint32_t bar_empty = smem_u32(bars + 1 + TC_MAX_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 1 + 2 * TC_MAX_STAGES);
  const uint32_t bar_tempty = smem_u32(bars + 3 + 2 * TC_MAX_STAGES);
  const uint32_t bar_qfree = smem_u32(bars + 5 + 2 * TC_MAX_STAGES);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 6 + 2 * TC_MAX_STAGES);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader
  unsigned long long dbg_c0 = 0, dbg_t0 = 0;
  if (p.debug && blockIdx.x == 0 && tid == 0) {
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }
  const uint32_t unit = PAIR ? blockIdx.x >> 1 : blockIdx.x;  // CTA or CTA pair

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_qfree, 1);
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (uint32_t a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, PAIR ? 2 * EW : EW);  // epilogue warps (of both CTAs) that drain an accumulator
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(smem_u32(tmem_ptr_s), TC_TMEM_COLS);
    else tmem_alloc(smem_u32(tmem_ptr_s), TC_TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers must exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // Work of this unit: segments (query tile, slots [s_begin, s_end)); slot -> tile is the identity in SCAN
  // mode and a stride over the whole corpus in DUMP mode.  The n_qt x n_es units of the regular grid have one
  // segment each (query tile group g, one of n_es shares of the first n_slots - n_tail slots).  When n_qt does
  // not divide the SM count the n_rem CTAs left over share the last n_tail slots of EVERY query tile: unit j
  // takes an equal run of the n_qt x n_tail tile jobs in query-tile-major order, i.e. a few query tiles one
  // after the other over (part of) the tail rows, reloading its resident query tile in between.
  const uint32_t n_main = p.n_slots - p.n_tail;
  auto segment = [&](uint32_t i, uint32_t& qt, uint32_t& s_begin, uint32_t& s_end) -> bool {
    const uint32_t n_full = p.n_qt * p.n_es;
    if (unit < n_full) {
      if (i) return false;
      const uint32_t g = unit % p.n_qt, es = unit / p.n_qt;
      qt = PAIR ? 2 * g + rank : g;
      s_begin = (uint32_t)(((uint64_t)es * n_main) / p.n_es);
      s_end = (uint32_t)(((uint64_t)(es + 1) * n_main) / p.n_es);
      return s_end > s_begin;
    }
    if (PAIR) return false;
    const uint64_t jobs = (uint64_t)p.n_qt * p.n_tail;
    const uint64_t lo = ((uint64_t)(unit - n_full) * jobs) / p.n_rem, hi = ((uint64_t)(unit - n_full + 1) * jobs) / p.n_rem;
    if (lo >= hi) return false;
    const uint64_t t = lo / p.n_tail + i;  // i-th query tile this run touches
    if (t * p.n_tail >= hi) return false;
    const uint64_t a = lo > t * p.n_tail ? lo : t * p.n_tail, b = hi < (t + 1) * p.n_tail ? hi : (t + 1) * p.n_tail;
    qt = (uint32_t)t;
    s_begin = n_main + (uint32_t)(a - t * p.n_tail);
    s_end = n_main + (uint32_t)(b - t * p.n_tail);
    return true;
  };
  auto tile_of = [&](uint32_t slot) {
    return p.mode == TC_MODE_SCAN ? p.tile0 + slot : (uint32_t)(((uint64_t)slot * p.n_tiles) / p.n_slots);
  };

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t qt, s_begin, s_end, it = 0;
      if (!PAIR) {
        for (uint32_t si = 0; segment(si, qt, s_begin, s_end); ++si) {
          if (QRES) {
            // a later segment replaces the resident query tile once the MMAs that read the previous one retired
            if (si) mbar_wait(bar_qfree, (si - 1) & 1);
            mbar_arrive_expect_tx(bar_q, p.n_kc * TC_QCHUNK_BYTES);
            for (uint32_t kc = 0; kc < p.n_kc; ++kc)
              tma_load_2d(smem_u32(sQ + (size_t)kc * TC_QCHUNK_BYTES), &tmQ, (int)(kc * TC_BK), (int)(qt * TC_BM),
                          bar_q);
          }
          for (uint32_t slot = s_begin; slot < s_end; ++slot) {
            const int row0 = (int)(tile_of(slot) * TC_BN);
            for (uint32_t kc = 0; kc < p.n_kc; ++kc, ++it) {
              const uint32_t stage = it % S;
              if (it >= S) mbar_wait(bar_empty + 8 * stage, ((it / S) - 1) & 1);
              if (p.debug & 4) {  // measurement: no E traffic at all, the MMAs reuse whatever the stage holds
                mbar_arrive(bar_full + 8 * stage);
                continue;
              }
              mbar_arrive_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
              tma_load_2d(smem_u32(sE + (size_t)stage * STAGE_BYTES), &tmE, (int)(kc * TC_BK), row0,
                          bar_full + 8 * stage);
              if (!QRES)
                tma_load_2d(smem_u32(sE + (size_t)stage * STAGE_BYTES + E_BYTES), &tmQ, (int)(kc * TC_BK),
                            (int)(qt * TC_BM), bar_full + 8 * stage);
            }
          }
        }
      } else if (segment(0, qt, s_begin, s_end)) {
        // both CTAs load; every transfer completes on the LEADER's barrier, which the leader
        // arms with the bytes of both halves.  A CTA reuses a stage when its own empty barrier
        // (signalled in both CTAs by the leader's commit) says the MMAs that read it retired.
        if (QRES) {
          const uint32_t l_bar_q = mapa_u32(bar_q, 0);
          if (rank == 0) mbar_arrive_expect_tx(bar_q, 2 * p.n_kc * TC_QCHUNK_BYTES);
          for (uint32_t kc = 0; kc < p.n_kc; ++kc)
            tma_load_2d_pair(smem_u32(sQ + (size_t)kc * TC_QCHUNK_BYTES), &tmQ, (int)(kc * TC_BK),
                             (int)(qt * TC_BM), l_bar_q);
        }
        for (uint32_t slot = s_begin; slot < s_end; ++slot) {
          const int row0 = (int)(tile_of(slot) * TC_BN + rank * (TC_BN / 2));
          for (uint32_t kc = 0; kc < p.n_kc; ++kc, ++it) {
            const uint32_t stage = it % S;
            if (it >= S) mbar_wait(bar_empty + 8 * stage, ((it / S) - 1) & 1);
            if (p.debug & 4) {
              if (rank == 0) mbar_arrive(bar_full + 8 * stage);
              continue;
            }
            if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * STAGE_BYTES);
            tma_load_2d_pair(smem_u32(sE + (size_t)stage * STAGE_BYTES), &tmE, (int)(kc * TC_BK), row0,
                             mapa_u32(bar_full + 8 * stage, 0));
            if (!QRES)
              tma_load_2d_pair(smem_u32(sE + (size_t)stage * STAGE_BYTES + E_BYTES), &tmQ, (int)(kc * TC_BK),
                               (int)(qt * TC_BM), mapa_u32(bar_full + 8 * stage, 0));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA of a pair) ----------
    if (lane == 0 && rank == 0) {
      uint32_t qt, s_begin, s_end, it = 0, tcount = 0;
      for (uint32_t si = 0; segment(si, qt, s_begin, s_end); ++si) {
        if (QRES) {
          mbar_wait(bar_q, si & 1);
          tc_fence_after();
        }
        for (uint32_t slot = s_begin; slot < s_end; ++slot, ++tcount) {
          const uint32_t acc = tcount & 1, use = tcount >> 1;
          if (use > 0) mbar_wait(bar_tempty + 8 * acc, (use - 1) & 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TC_BN;
          for (uint32_t kc = 0; kc < p.n_kc; ++kc, ++it) {
            const uint32_t stage = it % S;
            mbar_wait(bar_full + 8 * stage, (it / S) & 1);
            tc_fence_after();
            const uint64_t adesc = make_sw128_desc(
                smem_u32(QRES ? sQ + (size_t)kc * TC_QCHUNK_BYTES : sE + (size_t)stage * STAGE_BYTES + E_BYTES));
            const uint64_t bdesc = make_sw128_desc(smem_u32(sE + (size_t)stage * STAGE_BYTES));
#pragma unroll
            for (uint32_t k = 0; k < TC_BK / TC_UK; ++k) {
              // advance 16 bf16 = 32 B inside the swizzle atom: +2 in the (addr >> 4) field
              if (PAIR) umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, TC_IDESC_PAIR, (kc | k) != 0 ? 1u : 0u);
              else umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, TC_IDESC, (kc | k) != 0 ? 1u : 0u);
            }
            // frees the smem stage (in both CTAs) when these MMAs retire
            if (PAIR) umma_commit_pair(bar_empty + 8 * stage);
            else umma_commit(bar_empty + 8 * stage);
          }
          // accumulator complete (in both CTAs' TMEM)
          if (PAIR) umma_commit_pair(bar_tfull + 8 * acc);
          else umma_commit(bar_tfull + 8 * acc);
        }
        if (QRES && !PAIR) umma_commit(bar_qfree);  // the resident query tile may be replaced
      }
    }
  } else if (warp >= TC_CTRL_WARPS) {
    // ------------------------------ epilogue ----------------------------------
    // EW / 4 warps per TMEM lane quarter (warp w may read lanes 32 * (w % 4) ..); each takes an
    // equal share of the tile's 256 columns, 32 at a time
    constexpr uint32_t LIST_CAP = tc_list_cap(EW);
    constexpr uint32_t CH_PER_WARP = (TC_BN / 32) / (EW / 4);
    const uint32_t ew = warp & 3, part = (warp - TC_CTRL_WARPS) >> 2;
    const uint32_t list_slot = blockIdx.x * (EW * 32) + (warp - TC_CTRL_WARPS) * 32;
    const uint32_t l_bar_tempty = PAIR ? mapa_u32(bar_tempty, 0) : bar_tempty;  // the MMA issuer's barrier
    uint64_t* myL = p.lists + ((size_t)list_slot + lane) * LIST_CAP;
    uint32_t qt, s_begin, s_end, tcount = 0;
    for (uint32_t si = 0; segment(si, qt, s_begin, s_end); ++si) {
      const uint32_t q = qt * TC_BM + ew * 32 + lane;
      const bool valid = q < p.nq_valid;
      uint32_t cnt = 0;
      uint64_t tau_key = 0ull;
      float tau = valid ? -INFINITY : INFINITY;
      if (valid && p.mode == TC_MODE_SCAN) {
        tau_key = *((volatile uint64_t*)(p.gtau + q));
        if (tau_key != 0ull) tau = float_from_ord(key_ord(tau_key));
      }

      for (uint32_t slot = s_begin; slot < s_end; ++slot, ++tcount) {
        const uint32_t acc = tcount & 1, use = tcount >> 1;
        const uint32_t row_base = tile_of(slot) * TC_BN;
        mbar_wait(bar_tfull + 8 * acc, use & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((ew * 32u) << 16) + acc * TC_BN;
        if (p.debug & 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(l_bar_tempty + 8 * acc);
            else mbar_arrive(bar_tempty + 8 * acc);
          }
          continue;
        }
#pragma unroll 1
        for (uint32_t c = 0; c < CH_PER_WARP; ++c) {
          const uint32_t ch = part * CH_PER_WARP + c;
          float v[32];
          tmem_ld32(taddr + ch * 32, v);
          if (c == CH_PER_WARP - 1) {  // this warp's share of the accumulator is drained
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(l_bar_tempty + 8 * acc);
              else mbar_arrive(bar_tempty + 8 * acc);
            }
          }
          const uint32_t r0 = row_base + ch * 32;
          if (p.mode == TC_MODE_DUMP) {
            if (valid) {
              float* out = p.dump + (size_t)q * ((size_t)p.n_slots * TC_BN) + (size_t)slot * TC_BN + ch * 32;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o;
                float* po = reinterpret_cast<float*>(&o);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t row = r0 + j + u;
                  bool okr = row < p.n_rows;
                  if (okr && p.check_rows) okr = row_passes(p.flt, p.meta, p.agent, row);
                  po[u] = okr ? v[j + u] : -INFINITY;
                }
                *reinterpret_cast<float4*>(out + j) = o;
              }
            }
            continue;
          }
          // common path: the largest of the 32 scores against the cut-off (group maxima of 8 are
          // kept so that a hit is located without a per-score test; fmaxf drops NaN, padded
          // lanes have tau = +inf)
          float g8[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = fmaxf(fmaxf(v[8 * i], v[8 * i + 1]), v[8 * i + 2]);
            const float b = fmaxf(fmaxf(v[8 * i + 3], v[8 * i + 4]), v[8 * i + 5]);
            g8[i] = fmaxf(fmaxf(a, b), fmaxf(v[8 * i + 6], v[8 * i + 7]));
          }
          const bool hit = fmaxf(fmaxf(g8[0], g8[1]), fmaxf(g8[2], g8[3])) >= tau && !(p.debug & 2);
          if (__any_sync(0xffffffffu, hit)) {
            // make room first: a chunk can append up to 32 keys
            uint32_t full = __ballot_sync(0xffffffffu, cnt + 32 > LIST_CAP);
            if (p.static_tau) {
              if (cnt + 32 > LIST_CAP) {
                const uint32_t pos = atomicAdd(p.cnt + q, cnt);
                uint64_t* out = p.keys + (size_t)q * p.cap;
                for (uint32_t i = 0; i < cnt; ++i)
                  if (pos + i < p.cap) out[pos + i] = myL[i];
                cnt = 0;
              }
              full = 0;
            }
            while (full) {
              const uint32_t l = __ffs(full) - 1;
              full &= full - 1;
              __syncwarp();  // lane l's appended keys must be visible to the whole warp
              const uint32_t n_l = __shfl_sync(0xffffffffu, cnt, l);
              uint64_t kth =
                  warp_compact<(int)(LIST_CAP / 32)>(p.lists + ((size_t)list_slot + l) * LIST_CAP, n_l, p.KP, lane);
              if (lane == l) {
                cnt = n_l < p.KP ? n_l : p.KP;
                if (kth > tau_key) {
                  tau_key = kth;
                  tau = float_from_ord(key_ord(kth));
                  atomicMax(reinterpret_cast<unsigned long long*>(p.gtau + q), (unsigned long long)kth);
                }
              }
            }
            // hit mask (bit j: v[j] >= tau), then one loop iteration per hit; v[j] for a run-time j
            // comes from a 5-level select tree so v never leaves the registers
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) m |= (v[j] >= tau ? 1u : 0u) << j;
            if (!hit) m = 0;
            while (m) {
              const uint32_t j = __ffs(m) - 1;
              m &= m - 1;
              const float sc = pick32(v, j);
              const uint32_t row = r0 + j;
              if (sc >= tau && row < p.n_rows) {
                const uint64_t key = make_key(ord_from_float(sc), row);
                if (key >= tau_key && (!p.check_rows || row_passes(p.flt, p.meta, p.agent, row))) myL[cnt++] = key;
              }
            }
          }
        }
      }

      // final (SCAN): append what can still matter to the query's merged list
      if (p.mode == TC_MODE_SCAN && valid && cnt) {
        const uint64_t g = *((volatile uint64_t*)(p.gtau + q));
        uint32_t m = 0;
        for (uint32_t i = 0; i < cnt; ++i) m += myL[i] >= g;
        if (m) {
          uint32_t pos = atomicAdd(p.cnt + q, m);
          uint64_t* out = p.keys + (size_t)q * p.cap;
          for (uint32_t i = 0; i < cnt; ++i) {
            const uint64_t key = myL[i];
            if (key >= g) {
              if (pos < p.cap) out[pos] = key;
              ++pos;
            }
          }
        }
      }
      __syncwarp();  // the private lists are reused by the next segment
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // neither CTA may leave (or free TMEM) while the other still uses it
  else __syncthreads();
  if (p.debug && blockIdx.x == 0 && tid == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    g_tensor_clock[0] = clock64() - dbg_c0;
    g_tensor_clock[1] = t1 - dbg_t0;
  }
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TC_TMEM_COLS);
    else tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ---- cut-off from the sampled scores: tau0[q] = KP-th largest of dump[q][0..S) --------
// (NT threads per query: 1024 when few queries leave most SMs idle anyway -- the kernel is then a latency chain
// of block-wide passes over the sample -- 256 when there is a block per SM and more)
template <int NT>
__global__ void __launch_bounds__(NT) tau_select_kernel(const float* __restrict__ dump, uint32_t S, uint32_t KP,
                                                        uint64_t* __restrict__ gtau) {
  __shared__ uint32_t scratch[260];
  __shared__ uint32_t stage[4096];
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  const float* d = dump + (size_t)q * S;
  // ineligible entries (-inf, NaN) map to 0, below every real score: if fewer than KP sampled
  // rows are eligible the KP-th largest is 0 and the query simply gets no cut-off
  auto get = [&](uint32_t i) {
    const float x = d[i];
    return x > -INFINITY ? ord_from_float(x) : 0u;
  };
  const uint32_t t = S >= KP ? block_kth_largest(get, S, KP, scratch, stage, 4096u, tid, NT) : 0u;
  if (tid == 0) gtau[q] = (uint64_t)t << 32;  // the lowest key with that score (0 = none)
}

// ---- query preparation: normalise + convert to bf16, zero padded to [n_qt*128][ld16] --
__global__ void query_bf16_kernel(const float* __restrict__ Q, uint32_t ldq, uint32_t dim, uint32_t nq,
                                  uint32_t nq_pad, __nv_bfloat16* __restrict__ Q16, uint32_t ld16) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nq_pad) return;
  __nv_bfloat16* out = Q16 + (size_t)w * ld16;
  if (w >= nq) {
    for (uint32_t d = lane; d < ld16; d += 32) out[d] = __float2bfloat16_rn(0.0f);
    return;
  }
  const float* q = Q + (size_t)w * ldq;
  float ss = 0.0f;
  for (uint32_t d = lane; d < dim; d += 32) ss = fmaf(q[d], q[d], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = rsqrtf(ss);
  for (uint32_t d = lane; d < ld16; d += 32) out[d] = __float2bfloat16_rn(d < dim ? q[d] * inv : 0.0f);
}

void launch_query_bf16(const float* Q, uint32_t ldq, uint32_t dim, uint32_t nq, uint32_t nq_pad, void* Q16,
                       uint32_t ld16, cudaStream_t s) {
  if (!nq_pad) return;
  const uint32_t threads = 256, wpb = threads / 32;
  query_bf16_kernel<<<(nq_pad + wpb - 1) / wpb, threads, 0, s>>>(Q, ldq, dim, nq, nq_pad, (__nv_bfloat16*)Q16, ld16);
}

// ---- host side ---------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static bool encode_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the query tile stays resident when that leaves room for at least three ring stages
static bool tensor_q_resident(uint32_t n_kc) {
  return (size_t)n_kc * TC_QCHUNK_BYTES + 3 * tc_stage_bytes(false) + 1024 <= TC_SMEM_LIMIT;
}

static uint32_t tensor_stages(uint32_t n_kc, bool pair, size_t* total) {
  const bool qres = tensor_q_resident(n_kc);
  const size_t fixed = (qres ? (size_t)n_kc * TC_QCHUNK_BYTES : 0) + (6 + 2 * TC_MAX_STAGES) * 8 + 64;
  for (uint32_t s = TC_MAX_STAGES; s >= 2; --s) {
    size_t t = fixed + (size_t)s * (tc_stage_bytes(pair) + (qres ? 0 : TC_QCHUNK_BYTES));
    if (t <= TC_SMEM_LIMIT) {
      *total = t;
      return s;
    }
  }
  *total = 0;
  return 0;
}

// keys kept per query by the tensor pass for a request of k neighbours (0 = not served)
uint32_t tensor_keep(uint32_t k) {
  if (k + 16 <= 32) return 32;
  if (k + 16 <= 64) return 64;
  if (k + 16 <= 128) return 128;
  return 0;
}

bool tensor_scan_eligible(uint32_t ld16, uint32_t k) {
  size_t t;
  return get_encode() != nullptr && tensor_keep(k) != 0 && tensor_stages(ld16 / TC_BK, false, &t) != 0;
}

size_t tensor_scratch_bytes(int sm_count) { return (size_t)sm_count * TC_LIST_POOL * 8; }

// sampled row tiles for the cut-off bootstrap: more when few queries share the cost
uint32_t tensor_sample_tiles(uint32_t n_rows, uint32_t nq) {
  const uint32_t n_tiles = (n_rows + TC_BN - 1) / TC_BN;
  uint32_t want = nq <= 512 ? 64 : 32;  // 16384 / 8192 rows
  return want < n_tiles ? want : n_tiles;
}

template <int EW, bool QRES>
static cudaError_t launch_kernel(bool pair, uint32_t units, size_t smem, cudaStream_t s, const CUtensorMap& tmQ,
                                 const CUtensorMap& tmE, const TensorParams& p) {
  constexpr uint32_t threads = (TC_CTRL_WARPS + EW) * 32;
  if (!pair) {
    cudaError_t e = raise_dynamic_smem<tensor_scan_kernel<false, EW, QRES>>(smem);
    if (e != cudaSuccess) return e;
    tensor_scan_kernel<false, EW, QRES><<<units, threads, smem, s>>>(tmQ, tmE, p);
    return cudaGetLastError();
  }
  cudaError_t e = raise_dynamic_smem<tensor_scan_kernel<true, EW, QRES>>(smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * units, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, tensor_scan_kernel<true, EW, QRES>, tmQ, tmE, p);
}

// probe builds: report the effective SM clock of block 0 in the last debug-mode launch
void tensor_report_clock() {
  unsigned long long c[2] = {0, 0};
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(c, g_tensor_clock, sizeof c) == cudaSuccess && c[1])
    fprintf(stderr, "[cortex_gpu] tensor_scan block 0: %llu cycles in %llu ns = %.0f MHz\n", c[0], c[1],
            1e3 * (double)c[0] / (double)c[1]);
}

static cudaError_t launch_mode(const StoreView& st, const void* Q16, uint32_t q0, uint32_t nq, const DevFilter& flt,
                               bool check_rows, const CandView& cv, uint64_t* lists, float* dump, uint32_t n_slots,
                               uint32_t tile0, uint32_t mode, int sm_count, cudaStream_t s, const TensorTuning& tune,
                               bool static_tau = false) {
  const uint32_t n_tiles_q = (nq + TC_BM - 1) / TC_BM;
  if (n_tiles_q > (uint32_t)sm_count) return cudaErrorInvalidValue;  // caller splits larger batches
  // two or more query tiles: CTA pairs (each pair = two query tiles walking the same rows)
  const bool pair = tune.pair && n_tiles_q >= 2 && sm_count >= 2;
  uint32_t n_qt, n_es;  // query-tile groups (tiles, or pairs of tiles) x row splits
  if (pair) {
    n_qt = (n_tiles_q + 1) / 2;
    n_es = (uint32_t)(sm_count / 2) / n_qt;
  } else {
    n_qt = n_tiles_q;
    n_es = (uint32_t)sm_count / n_qt;
  }
  if (n_es < 1) n_es = 1;
  TensorParams p;
  p.n_rows = st.n_rows;
  p.n_tiles = (st.n_rows + TC_BN - 1) / TC_BN;
  p.n_slots = n_slots;  // SCAN: tiles [tile0, tile0 + n_slots); DUMP: sampled tiles
  p.tile0 = tile0;
  if (mode == TC_MODE_SCAN && (uint64_t)tile0 + n_slots > p.n_tiles) return cudaErrorInvalidValue;
  if (n_es > p.n_slots) n_es = p.n_slots;
  p.n_kc = st.ld16 / TC_BK;
  p.n_qt = n_qt;
  p.n_es = n_es;
  // SMs left over by the n_qt x n_es grid (8 query tiles on 148 SMs: 4) take an equal share of the work from the
  // end of the slot range, every query tile of it (see the kernel)
  p.n_tail = 0;
  p.n_rem = 0;
  if (!pair && tune.use_leftover_sms) {
    const uint32_t n_full = n_qt * n_es;
    const uint32_t rem = (uint32_t)sm_count > n_full ? (uint32_t)sm_count - n_full : 0u;
    // ... when every left-over CTA then gets WHOLE query tiles: they walk the tail rows in step (one HBM read,
    // the later walks from L2) like the CTAs of the regular grid; uneven shares would scatter them over the tail
    if (rem && n_qt % rem == 0 && p.n_slots >= 2u * (uint32_t)sm_count) {
      p.n_tail = (uint32_t)(((uint64_t)p.n_slots * rem) / (n_full + rem));
      if (p.n_tail) p.n_rem = rem;
    }
  }
  const uint32_t units = n_qt * n_es + p.n_rem;
  p.nq_valid = nq;
  size_t smem;
  p.stages = tensor_stages(p.n_kc, pair, &smem);
  if (!p.stages) return cudaErrorInvalidConfiguration;
  p.mode = mode;
  p.check_rows = check_rows ? 1u : 0u;
#ifdef CX_PROBE
  p.debug = mode == TC_MODE_SCAN ? (uint32_t)tune.debug : 0u;
#else
  p.debug = 0u;  // the measurement hook exists in probe builds only
#endif
  p.static_tau = static_tau ? 1u : 0u;
  p.KP = cv.KP;
  p.meta = st.meta;
  p.agent = st.agent;
  p.flt = flt;
  p.lists = lists;
  p.keys = cv.keys + (size_t)(q0 - cv.q_base) * cv.cap;
  p.cnt = cv.cnt + q0;
  p.gtau = cv.gtau + q0;
  p.cap = cv.cap;
  p.dump = dump;
  CUtensorMap tmQ, tmE;
  const __nv_bfloat16* qbase = (const __nv_bfloat16*)Q16 + (size_t)q0 * st.ld16;
  // the query buffer is padded to whole 128-row tiles; a pair's missing second tile is out of
  // bounds for the map and reads as zeros
  if (!encode_2d(&tmQ, qbase, st.ld16, (uint64_t)n_tiles_q * TC_BM, TC_BM)) return cudaErrorInvalidValue;
  if (!encode_2d(&tmE, st.E16, st.ld16, st.n_rows, pair ? TC_BN / 2 : TC_BN)) return cudaErrorInvalidValue;
  if (!tensor_q_resident(p.n_kc)) return launch_kernel<8, false>(pair, units, smem, s, tmQ, tmE, p);
  return tune.epi_warps == 8 ? launch_kernel<8, true>(pair, units, smem, s, tmQ, tmE, p)
                             : launch_kernel<16, true>(pair, units, smem, s, tmQ, tmE, p);
}

// Bootstrap: sample scores -> per-query cut-off in cv.gtau[q0 .. q0+nq).  dump holds
// nq * n_slots * 256 floats.
cudaError_t launch_tensor_bootstrap(const StoreView& st, const void* Q16, uint32_t q0, uint32_t nq,
                                    const DevFilter& flt, bool check_rows, const CandView& cv, float* dump,
                                    uint32_t n_slots, int sm_count, cudaStream_t s, const TensorTuning& tune) {
  if (!nq || !st.n_rows || !n_slots) return cudaSuccess;
  cudaError_t e =
      launch_mode(st, Q16, q0, nq, flt, check_rows, cv, nullptr, dump, n_slots, 0, TC_MODE_DUMP, sm_count, s, tune);
  if (e != cudaSuccess) return e;
  if (nq <= (uint32_t)sm_count) tau_select_kernel<1024><<<nq, 1024, 0, s>>>(dump, n_slots * TC_BN, cv.KP, cv.gtau + q0);
  else tau_select_kernel<256><<<nq, 256, 0, s>>>(dump, n_slots * TC_BN, cv.KP, cv.gtau + q0);
  return cudaGetLastError();
}

cudaError_t launch_tensor_scan(const StoreView& st, const void* Q16, uint32_t q0, uint32_t nq,
                               const DevFilter& flt, bool check_rows, const CandView& cv, uint64_t* lists,
                               uint32_t tile0, uint32_t n_tiles, int sm_count, cudaStream_t s, const TensorTuning& tune,
                               bool static_tau) {
  if (!nq || !st.n_rows || !n_tiles) return cudaSuccess;
  return launch_mode(st, Q16, q0, nq, flt, check_rows, cv, lists, nullptr, n_tiles, tile0, TC_MODE_SCAN, sm_count, s,
                     tune, static_tau);
}

// gtau[q] = the largest key below every key whose approximate cosine is thr_cos or more
__global__ void fill_tau_kernel(uint64_t* __restrict__ gtau, uint32_t nq, float thr_cos) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) gtau[i] = ((uint64_t)ord_from_float(thr_cos) << 32);
}
void launch_fill_tau(const CandView& cv, uint32_t q0, uint32_t nq, float thr_cos, cudaStream_t s) {
  if (nq) fill_tau_kernel<<<(nq + 255) / 256, 256, 0, s>>>(cv.gtau + q0, nq, thr_cos);
}

uint32_t tensor_tiles(uint32_t n_rows) { return (n_rows + TC_BN - 1) / TC_BN; }

// ---- cut-off refinement between two phases of a scan -----------------------------------
// The merged list of a query holds every row scanned so far whose score cleared the cut-off in
// force when it was seen; the KP-th best of them is a (much tighter) lower bound of the KP-th
// best over the whole corpus, so the next phase nominates far fewer rows.
// `margin` (cosine units, twice the pass's error bound) is subtracted so that the list keeps every row
// inside the +-eps band around the k-th result, which the select kernel may have to rescore.
__global__ void __launch_bounds__(256) tau_refine_kernel(const uint64_t* __restrict__ keys,
                                                         const uint32_t* __restrict__ cnt, uint32_t cap, uint32_t KP,
                                                         float margin, uint64_t* __restrict__ gtau) {
  __shared__ uint32_t scratch[260];
  __shared__ uint32_t stage[4096];
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  const uint32_t n = min(cnt[q], cap);  // an overflowed list still holds `cap` real rows
  if (n < KP) return;
  const uint64_t* k = keys + (size_t)q * cap;
  auto get = [&](uint32_t i) { return key_ord(k[i]); };
  const uint32_t t = block_kth_largest(get, n, KP, scratch, stage, 4096u, tid, 256);
  if (tid == 0) {
    const uint64_t nt = (uint64_t)ord_from_float(float_from_ord(t) - margin) << 32;  // the lowest key with that score
    if (nt > gtau[q]) gtau[q] = nt;
  }
}

cudaError_t launch_tau_refine(const CandView& cv, uint32_t q0, uint32_t nq, float margin, cudaStream_t s) {
  if (!nq) return cudaSuccess;
  tau_refine_kernel<<<nq, 256, 0, s>>>(cv.keys + (size_t)(q0 - cv.q_base) * cv.cap, cv.cnt + q0, cv.cap, cv.KP,
                                       margin, cv.gtau + q0);
  return cudaGetLastError();
}

}  // namespace cx
