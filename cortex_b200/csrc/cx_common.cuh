// cx_common.cuh -- shared device helpers for the cortex_b200 similarity scan.
//
// Arithmetic contract: every score that leaves the library is
// produced by ref_* below, which restates
//   /root/reference/crates/cortex-core/src/vector/index.rs:169-179 (distance)
//   /root/reference/crates/cortex-core/src/vector/index.rs:253-256 (clamp)
// with round-to-nearest intrinsics that ptxas may not contract or reassociate.
// The fast passes (streaming fp32, tcgen05 bf16) only nominate candidates.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cx {

// ---- candidate keys ---------------------------------------------------------
// 64-bit key: high word orders by score, low word by (0xFFFFFFFF - row) so that
// among equal scores the lower row (earlier insertion) is the larger key.
// Key 0 is "empty"; every real key is > 0.

__host__ __device__ __forceinline__ uint32_t ord_from_float(float f) {
  // order-preserving map of a non-NaN float to uint32 (negatives below positives)
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float float_from_ord(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t ord, uint32_t row) {
  return ((uint64_t)ord << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }
__host__ __device__ __forceinline__ uint32_t key_ord(uint64_t k) { return (uint32_t)(k >> 32); }

// Final-score ordering: scores are clamped to [0,1] or NaN.  NaN sorts last
// (ord 1), -0.0 ties with +0.0 (partial_cmp says Equal), others by value.
__host__ __device__ __forceinline__ uint32_t ord_from_score(float s) {
  if (s != s) return 1u;
  if (s == 0.0f) s = 0.0f;  // folds -0.0
  return ord_from_float(s);  // >= 0x80000000
}

#ifdef __CUDACC__
// ---- reference arithmetic -----------------------------------------------------
// index.rs:172-174: Iterator::sum::<f32>() is a strict left fold; products are
// rounded before the add (Rust never contracts to FMA).
__device__ __forceinline__ float ref_fold(float acc, float a, float b) {
  return __fadd_rn(acc, __fmul_rn(a, b));
}
// index.rs:176-177
__device__ __forceinline__ float ref_distance_from(float dot, float norm_a, float norm_b) {
  float sim = __fdiv_rn(dot, __fmul_rn(norm_a, norm_b));
  return __fsub_rn(1.0f, sim);
}
// index.rs:254-256 with Rust f32::clamp semantics (NaN stays, -0.0 stays)
__device__ __forceinline__ float ref_score_from_distance(float d) {
  float s = __fsub_rn(1.0f, d);
  if (s < 0.0f) s = 0.0f;
  if (s > 1.0f) s = 1.0f;
  return s;
}

// norms outside this range come from squares that under- / overflowed fp32: the reference's
// arithmetic is followed on the exact path only (rows: cx_exact.cu row_norm_kernel; queries: the
// rescoring kernels report "not verified")
constexpr float NORM_REGULAR_MIN = 1e-15f;
constexpr float NORM_REGULAR_MAX = 1e15f;

// ---- per-row metadata word ---------------------------------------------------
constexpr uint32_t META_DEAD = 0x80000000u;
constexpr uint32_t META_HAS = 0x40000000u;
constexpr uint32_t META_KIND_MASK = 0xFFu;
constexpr uint32_t AGENT_NONE = 0xFFFFFFFFu;

// Device form of VectorFilter (index.rs:17-26) + matches_filter (index.rs:225-251)
struct DevFilter {
  uint64_t kind_mask[4];     // bit per interned kind id (<= 255 kinds)
  const uint32_t* excl_rows; // rows of the excluded ids that exist in the index
  uint32_t n_excl;
  uint32_t agent;            // interned agent id or AGENT_NONE (matches nothing)
  int32_t has_kinds;
  int32_t has_agent;
};

__device__ __forceinline__ bool row_passes(const DevFilter& f, const uint32_t* __restrict__ meta,
                                           const uint32_t* __restrict__ agent, uint32_t row) {
  uint32_t m = __ldg(meta + row);
  if (m & META_DEAD) return false;
  for (uint32_t i = 0; i < f.n_excl; ++i)
    if (__ldg(f.excl_rows + i) == row) return false;
  if (m & META_HAS) {
    if (f.has_kinds) {
      uint32_t k = m & META_KIND_MASK;
      if (!((f.kind_mask[k >> 6] >> (k & 63)) & 1ull)) return false;
    }
    if (f.has_agent && __ldg(agent + row) != f.agent) return false;
  }
  return true;
}

// ---- PTX wrappers: mbarrier + 1-D bulk async copy (TMA engine, UBLKCP) --------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy completing on an mbarrier; 16 B aligned, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- block-wide bitonic sort (descending) of n = 2^m u64 keys in shared memory
// `sync` is called between stages by every participating thread.
template <typename SyncFn>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* a, uint32_t n, uint32_t tid, uint32_t nthreads,
                                                  SyncFn sync) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = tid; i < n; i += nthreads) {
        uint32_t l = i ^ j;
        if (l > i) {
          uint64_t x = a[i], y = a[l];
          bool desc_block = (i & k) == 0;
          if (desc_block ? (x < y) : (x > y)) {
            a[i] = y;
            a[l] = x;
          }
        }
      }
      sync();
    }
  }
}
// ---- block-wide k-th largest (1-based) of n uint32 values: MSB-first radix select ----
// get(i) returns value i (typically from global memory).  After the first (top byte)
// pass the values that share the selected byte are compacted into `stage` (shared
// memory, cap_stage words) so the remaining three passes do not touch global memory
// again; if they do not fit, those passes re-read get().  scratch: 260 words of shared
// memory.  All threads of the block call it (it synchronises with __syncthreads);
// requires 1 <= k <= n.
__device__ __forceinline__ void radix_pick_digit(uint32_t* scratch, uint32_t k, uint32_t tid) {
  // scratch[0..255] = histogram; writes scratch[256] = digit, scratch[257] = rank inside it
  if (tid < 32) {  // lane l owns bins [8l, 8l+8); suffix sums over lanes pick the digit
    uint32_t s = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) s += scratch[8 * tid + b];
    uint32_t suf = s;  // becomes the sum over lanes >= tid
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
      if (tid + o < 32) suf += t;
    }
    const uint32_t above = suf - s;  // sum over lanes > tid
    if (suf >= k && above < k) {
      uint32_t acc = above;
      int d = 7;
      for (; d > 0; --d) {
        if (acc + scratch[8 * tid + d] >= k) break;
        acc += scratch[8 * tid + d];
      }
      scratch[256] = 8 * tid + (uint32_t)d;
      scratch[257] = k - acc;
    }
  }
}

template <typename GetFn>
__device__ __forceinline__ uint32_t block_kth_largest(GetFn get, uint32_t n, uint32_t k, uint32_t* scratch,
                                                      uint32_t* stage, uint32_t cap_stage, uint32_t tid,
                                                      uint32_t nthreads) {
  // pass over the top byte
  for (uint32_t i = tid; i < 256; i += nthreads) scratch[i] = 0;
  if (tid == 0) scratch[258] = 0;
  __syncthreads();
  for (uint32_t i = tid; i < n; i += nthreads) atomicAdd(&scratch[get(i) >> 24], 1u);
  __syncthreads();
  radix_pick_digit(scratch, k, tid);
  __syncthreads();
  uint32_t prefix = scratch[256] << 24, mask = 0xFF000000u;
  k = scratch[257];
  // compact the survivors into shared memory
  for (uint32_t i = tid; i < n; i += nthreads) {
    const uint32_t v = get(i);
    if ((v & mask) == prefix) {
      const uint32_t pos = atomicAdd(&scratch[258], 1u);
      if (pos < cap_stage) stage[pos] = v;
    }
  }
  __syncthreads();
  const uint32_t n_st = scratch[258];
  const bool staged = n_st <= cap_stage;
#pragma unroll 1
  for (int pass = 2; pass >= 0; --pass) {
    const uint32_t shift = 8u * (uint32_t)pass;
    for (uint32_t i = tid; i < 256; i += nthreads) scratch[i] = 0;
    __syncthreads();
    if (staged) {
      for (uint32_t i = tid; i < n_st; i += nthreads) {
        const uint32_t v = stage[i];
        if ((v & mask) == prefix) atomicAdd(&scratch[(v >> shift) & 255u], 1u);
      }
    } else {
      for (uint32_t i = tid; i < n; i += nthreads) {
        const uint32_t v = get(i);
        if ((v & mask) == prefix) atomicAdd(&scratch[(v >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    radix_pick_digit(scratch, k, tid);
    __syncthreads();
    prefix |= scratch[256] << shift;
    mask |= 255u << shift;
    k = scratch[257];
    __syncthreads();
  }
  return prefix;
}
#endif  // __CUDACC__

}  // namespace cx
