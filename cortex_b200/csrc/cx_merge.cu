// cx_merge.cu -- the exchange step of the row-sharded search:
//   pack   : a rank's local top-k (shard-local rows, exact scores) -> fixed-size payload
//   merge  : the all-gathered payloads of W ranks -> global top-k per query
// Order = the single-index order: score descending, NaN last, then GLOBAL row ascending
// (shards are contiguous blocks in insertion order), i.e. the stable descending sort of
// vector/index.rs:287-292 with row order as tie order.
#include "cx_index.h"

namespace cx {

// payload per (query, slot): word0 = score-order key (high 32) | distance bits (low 32),
//                            word1 = global row; word0 == 0 marks an empty slot.
// Two trailer words follow the B*k slots: [0] = how many of this rank's queries are still unverified
// (their slots may change once the rank has retried them), [1] = reserved.  The trailer lets every
// rank learn from the gathered payloads alone whether the exchange has to be repeated.
__global__ void pack_topk_kernel(const uint32_t* __restrict__ rows, const float* __restrict__ score,
                                 const float* __restrict__ dist, const uint32_t* __restrict__ n,
                                 const uint32_t* __restrict__ ok, uint32_t B, uint32_t k, uint64_t row_offset,
                                 uint64_t* __restrict__ payload) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * k) return;
  if (ok && i < B && !ok[i]) atomicAdd(reinterpret_cast<unsigned long long*>(payload + 2 * (size_t)B * k), 1ull);
  const uint32_t b = i / k, j = i % k;
  uint64_t w0 = 0, w1 = 0;
  if (j < n[b]) {
    w0 = ((uint64_t)ord_from_score(score[i]) << 32) | (uint64_t)__float_as_uint(dist[i]);
    w1 = row_offset + rows[i];
  }
  payload[2 * (size_t)i] = w0;
  payload[2 * (size_t)i + 1] = w1;
}

// Does slot (x0, x1) come before (ord, row) in the merged order?  (empty slots never do)
__device__ __forceinline__ bool slot_before(uint64_t x0, uint64_t x1, uint32_t ord, uint64_t row) {
  if (x0 == 0) return false;
  const uint32_t xo = (uint32_t)(x0 >> 32);
  return xo > ord || (xo == ord && x1 < row);
}

// One CTA per query.  Every rank's list is already in merged order (score desc, NaN last, row asc) with its
// valid slots first, so the rank of a slot is its own position plus, for every other rank, the number of that
// rank's slots that come before it: one binary search per (slot, other rank) -- O(W k W log k) per query
// instead of the (W k)^2 of rank counting, which matters at k = 100 on 8 GPUs.  No shared memory, any k.
__global__ void merge_topk_kernel(const uint64_t* __restrict__ gathered, uint32_t W, uint32_t B, uint32_t k,
                                  int64_t* __restrict__ out_rows, float* __restrict__ out_score,
                                  float* __restrict__ out_dist, uint32_t* __restrict__ out_n,
                                  uint64_t* __restrict__ out_unverified) {
  const uint32_t b = blockIdx.x, tid = threadIdx.x, n = W * k;
  const size_t stride = 2 * (size_t)B * k + 2;  // words per rank: slots + trailer
  __shared__ uint32_t s_valid;
  if (tid == 0) s_valid = 0;
  if (b == 0 && tid == 0 && out_unverified) {
    uint64_t t = 0;
    for (uint32_t w = 0; w < W; ++w) t += gathered[(size_t)w * stride + stride - 2];
    *out_unverified = t;
  }
  __syncthreads();
  uint32_t valid = 0;
  for (uint32_t i = tid; i < n; i += blockDim.x) {
    const uint32_t w = i / k, j = i % k;
    const uint64_t* src = gathered + (size_t)w * stride + 2 * ((size_t)b * k + j);
    const uint64_t w0 = src[0], w1 = src[1];
    if (w0 == 0) continue;
    ++valid;
    const uint32_t ord = (uint32_t)(w0 >> 32);
    uint32_t rank = j;
    for (uint32_t x = 0; x < W; ++x) {
      if (x == w) continue;
      const uint64_t* L = gathered + (size_t)x * stride + 2 * (size_t)b * k;
      uint32_t lo = 0, hi = k;  // first slot of rank x that does NOT come before mine
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (slot_before(L[2 * mid], L[2 * mid + 1], ord, w1)) lo = mid + 1;
        else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      const size_t o = (size_t)b * k + rank;
      out_rows[o] = (int64_t)w1;
      out_dist[o] = __uint_as_float((uint32_t)w0);
      // the score is re-derived from the distance with the reference's own two operations
      out_score[o] = ref_score_from_distance(__uint_as_float((uint32_t)w0));
    }
  }
  if (valid) atomicAdd(&s_valid, valid);
  __syncthreads();
  if (tid == 0) out_n[b] = s_valid < k ? s_valid : k;
}

// ---- auto-link candidate post-pass ------------------------------------------------------
// Per new node b: walk its k results in order (best first), skip the node itself, keep
// score >= threshold (SimilarityLinkRule, linker/rules.rs:50), stop at max_edges
// (linker/auto_linker.rs:261).  One thread per node; k is ~100.
__global__ void autolink_filter_kernel(const uint32_t* __restrict__ rows, const float* __restrict__ score,
                                       const uint32_t* __restrict__ n, const uint32_t* __restrict__ self_rows,
                                       const uint8_t* __restrict__ ids, uint32_t B, uint32_t k, float threshold,
                                       uint32_t max_edges, uint32_t* __restrict__ out_rows,
                                       float* __restrict__ out_score, uint8_t* __restrict__ out_ids,
                                       uint32_t* __restrict__ out_n) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t self = self_rows ? self_rows[b] : 0xFFFFFFFFu;
  const uint32_t nb = n[b] < k ? n[b] : k;
  uint32_t m = 0;
  for (uint32_t j = 0; j < nb && m < max_edges; ++j) {
    const uint32_t r = rows[(size_t)b * k + j];
    const float s = score[(size_t)b * k + j];
    if (r == self) continue;
    if (!(s >= threshold)) continue;  // false for NaN
    const size_t o = (size_t)b * max_edges + m;
    out_rows[o] = r;
    out_score[o] = s;
    if (out_ids)
      *reinterpret_cast<uint4*>(out_ids + o * 16) = *reinterpret_cast<const uint4*>(ids + (size_t)r * 16);
    ++m;
  }
  out_n[b] = m;
}

// the same post-pass on MERGED lists of a row-sharded search (global rows, int64; self < 0 = not in the index)
__global__ void autolink_filter_global_kernel(const int64_t* __restrict__ rows, const float* __restrict__ score,
                                              const uint32_t* __restrict__ n, const int64_t* __restrict__ self_rows,
                                              uint32_t B, uint32_t k, float threshold, uint32_t max_edges,
                                              int64_t* __restrict__ out_rows, float* __restrict__ out_score,
                                              uint32_t* __restrict__ out_n) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t self = self_rows ? self_rows[b] : -1;
  const uint32_t nb = n[b] < k ? n[b] : k;
  uint32_t m = 0;
  for (uint32_t j = 0; j < nb && m < max_edges; ++j) {
    const int64_t r = rows[(size_t)b * k + j];
    const float s = score[(size_t)b * k + j];
    if (r == self) continue;
    if (!(s >= threshold)) continue;  // false for NaN
    const size_t o = (size_t)b * max_edges + m;
    out_rows[o] = r;
    out_score[o] = s;
    ++m;
  }
  out_n[b] = m;
}

void launch_autolink_filter(const uint32_t* rows, const float* score, const uint32_t* n, const uint32_t* self_rows,
                            const uint8_t* ids, uint32_t B, uint32_t k, float threshold, uint32_t max_edges,
                            uint32_t* out_rows, float* out_score, uint8_t* out_ids, uint32_t* out_n,
                            cudaStream_t s) {
  if (!B) return;
  autolink_filter_kernel<<<(B + 127) / 128, 128, 0, s>>>(rows, score, n, self_rows, ids, B, k, threshold, max_edges,
                                                         out_rows, out_score, out_ids, out_n);
}

__global__ void iota_kernel(uint32_t* __restrict__ dst, uint32_t n, uint32_t base) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = base + i;
}
void launch_iota(uint32_t* dst, uint32_t n, uint32_t base, cudaStream_t s) {
  if (n) iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(dst, n, base);
}

// ---- dedup scan: per-node partner lists -> dense triples ------------------------------------
// One CTA: exclusive prefix sum of the per-node counts (removed nodes count nothing: dedup.rs:73-76).
constexpr int CP_THREADS = 1024;
__global__ void __launch_bounds__(CP_THREADS) pair_offsets_kernel(uint32_t* __restrict__ n, uint32_t* __restrict__ tot,
                                                                  const uint32_t* __restrict__ meta, uint32_t r0,
                                                                  uint32_t B, uint32_t* __restrict__ off) {
  __shared__ uint32_t warp_sum[CP_THREADS / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t per = (B + CP_THREADS - 1) / CP_THREADS;
  const uint32_t i0 = tid * per, i1 = min(B, i0 + per);
  uint32_t mine = 0;
  for (uint32_t i = i0; i < i1; ++i) {
    if (meta[r0 + i] & META_DEAD) {
      n[i] = 0;
      if (tot) tot[i] = 0;
    }
    mine += n[i];
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_sum[lane];
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= (uint32_t)o) wi += t;
    }
    warp_sum[lane] = wi - w;  // exclusive over warps
  }
  __syncthreads();
  uint32_t run = warp_sum[warp] + incl - mine;
  for (uint32_t i = i0; i < i1; ++i) {
    off[i] = run;
    run += n[i];
  }
  if (tid == CP_THREADS - 1) off[B] = run;  // the last thread's chunk ends the array (empty chunks included)
}

// one warp per node: its partners, in order, to their dense positions
__global__ void pair_write_kernel(const uint32_t* __restrict__ rows, const float* __restrict__ score,
                                  const uint32_t* __restrict__ n, const uint32_t* __restrict__ off, uint32_t r0,
                                  uint32_t B, uint32_t kd, uint32_t* __restrict__ pairs, uint32_t limit) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B) return;
  const uint32_t m = n[w], base = off[w];
  for (uint32_t j = lane; j < m; j += 32) {
    const uint32_t pos = base + j;
    if (pos >= limit) break;
    pairs[3 * (size_t)pos] = r0 + w;
    pairs[3 * (size_t)pos + 1] = rows[(size_t)w * kd + j];
    pairs[3 * (size_t)pos + 2] = __float_as_uint(score[(size_t)w * kd + j]);
  }
}

void launch_compact_pairs(const uint32_t* rows, const float* score, uint32_t* n, uint32_t* tot, const uint32_t* meta,
                          uint32_t r0, uint32_t B, uint32_t kd, uint32_t* off, uint32_t* pairs, uint32_t limit,
                          cudaStream_t s) {
  if (!B) return;
  pair_offsets_kernel<<<1, CP_THREADS, 0, s>>>(n, tot, meta, r0, B, off);
  if (limit) pair_write_kernel<<<(B + 7) / 8, 256, 0, s>>>(rows, score, n, off, r0, B, kd, pairs, limit);
}

// the same prefix sum over counts with an explicit "removed" flag per node (multi-device dedup scan)
void launch_compact_offsets(uint32_t* n, const uint32_t* dead, uint32_t B, uint32_t* off, cudaStream_t s) {
  if (B) pair_offsets_kernel<<<1, CP_THREADS, 0, s>>>(n, nullptr, dead, 0, B, off);
}

}  // namespace cx

using namespace cx;

extern "C" cx_status cx_pack_topk_device(const uint32_t* d_rows, const float* d_score, const float* d_distance,
                                         const uint32_t* d_n, const uint32_t* d_ok, uint64_t B, uint64_t k,
                                         uint64_t row_offset, uint64_t* d_payload, void* stream) {
  cx::CallerDevice keep_callers_device;
  if (!d_rows || !d_score || !d_distance || !d_n || !d_payload) return fail(CX_ERR_VALIDATION, "null device buffer");
  if (!B || !k) return CX_OK;
  const uint64_t total = B * k;
  CU(cudaMemsetAsync(d_payload + 2 * total, 0, 16, (cudaStream_t)stream));  // trailer
  pack_topk_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_rows, d_score, d_distance, d_n, d_ok, (uint32_t)B, (uint32_t)k, row_offset, d_payload);
  CU(cudaGetLastError());
  return CX_OK;
}

extern "C" cx_status cx_merge_topk_device(const uint64_t* d_gathered, uint32_t world, uint64_t B, uint64_t k,
                                          int64_t* d_out_rows, float* d_out_score, float* d_out_distance,
                                          uint32_t* d_out_n, uint64_t* d_out_unverified, void* stream) {
  cx::CallerDevice keep_callers_device;
  if (!d_gathered || !d_out_rows || !d_out_score || !d_out_distance || !d_out_n)
    return fail(CX_ERR_VALIDATION, "null device buffer");
  if (!B || !k || !world) return CX_OK;
  const unsigned threads = world * k >= 256 ? 256u : 128u;
  merge_topk_kernel<<<(unsigned)B, threads, 0, (cudaStream_t)stream>>>(d_gathered, world, (uint32_t)B, (uint32_t)k,
                                                                      d_out_rows, d_out_score, d_out_distance, d_out_n,
                                                                      d_out_unverified);
  CU(cudaGetLastError());
  return CX_OK;
}

// AutoLinker::run_cycle's candidate post-pass (linker/auto_linker.rs:224-264, rules.rs:42-62) on the merged
// lists of a row-sharded search: skip self, keep score >= threshold, at most max_edges per node.
extern "C" cx_status cx_autolink_filter_device(const int64_t* d_rows, const float* d_score, const uint32_t* d_n,
                                               const int64_t* d_self_rows, uint64_t B, uint64_t k, float threshold,
                                               uint32_t max_edges_per_node, int64_t* d_out_rows, float* d_out_score,
                                               uint32_t* d_out_n, void* stream) {
  cx::CallerDevice keep_callers_device;
  if (!d_rows || !d_score || !d_n || !d_out_rows || !d_out_score || !d_out_n)
    return fail(CX_ERR_VALIDATION, "null device buffer");
  if (!B) return CX_OK;
  autolink_filter_global_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      d_rows, d_score, d_n, d_self_rows, (uint32_t)B, (uint32_t)k, threshold, max_edges_per_node, d_out_rows, d_out_score,
      d_out_n);
  CU(cudaGetLastError());
  return CX_OK;
}
