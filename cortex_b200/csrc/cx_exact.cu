// cx_exact.cu -- K3/K4 and the exact path: every number produced here follows the
// reference arithmetic bit for bit (see cx_common.cuh).
//
//   prepare_rows    : stored-row norms in reference order (index.rs:174 hoisted out
//                     of the per-pair loop; the value is identical because the row
//                     never changes), reciprocal norms and the bf16 shadow matrix
//   prepare_queries : query norm (index.rs:173), once per query instead of per pair
//   exact_keys      : full scan with reference arithmetic (brute_force_search,
//                     index.rs:259-294) producing sortable keys
//   exact_emit      : (row -> id, score, distance) for the head of a sorted key list
#include <cub/device/device_radix_sort.cuh>
#include <cuda_bf16.h>

#include "cx_kernels.h"

namespace cx {

// ---------------------------------------------------------------------------------
__global__ void row_norm_kernel(float* __restrict__ E, float* __restrict__ norm, float* __restrict__ rnorm,
                                uint32_t dim, uint32_t ld, uint32_t r0, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* row = E + (size_t)(r0 + i) * ld;
  float acc = 0.0f;
  bool any = false;
  uint32_t d = 0;
  for (; d + 4 <= dim; d += 4) {
    float4 v = *reinterpret_cast<const float4*>(row + d);
    acc = ref_fold(acc, v.x, v.x);
    acc = ref_fold(acc, v.y, v.y);
    acc = ref_fold(acc, v.z, v.z);
    acc = ref_fold(acc, v.w, v.w);
    any |= (v.x != 0.0f) | (v.y != 0.0f) | (v.z != 0.0f) | (v.w != 0.0f);
  }
  for (; d < dim; ++d) {
    acc = ref_fold(acc, row[d], row[d]);
    any |= row[d] != 0.0f;
  }
  for (d = dim; d < ld; ++d) row[d] = 0.0f;  // padding never contributes
  float nb = __fsqrt_rn(acc);
  norm[r0 + i] = nb;
  // A row whose squares under- or overflow fp32 has a reference norm that is not its length, so the
  // fast passes' error bounds do not hold for it (e.g. |x| ~ 1e-25: norm 0, sim = dot / 0 = +-inf,
  // score 1 or 0).  Such rows get a NaN reciprocal norm -- no fast pass ever nominates them -- and are
  // counted; an index that holds any is served by the exact path (cx_index.cu).  An all-zero row is
  // regular: its score is NaN for every query in the reference too.
  const bool irregular = any && !(nb >= NORM_REGULAR_MIN && nb <= NORM_REGULAR_MAX);
  rnorm[r0 + i] = irregular ? __int_as_float(0x7fc00000) : __frcp_rn(nb);
}

__global__ void count_nan_kernel(const float* __restrict__ x, const uint32_t* __restrict__ meta, uint32_t r0,
                                 uint32_t n, int sign, uint32_t* __restrict__ out) {
  uint32_t c = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    c += (x[r0 + i] != x[r0 + i]) && !(meta[r0 + i] & META_DEAD);
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) {
    if (sign > 0) atomicAdd(out, c);
    else atomicSub(out, c);
  }
}

// live rows of [r0, r0+n) with a NaN reciprocal norm (irregular rows, see row_norm_kernel) are added to
// (sign > 0) or taken off (sign < 0) the index's running count: the count follows inserts, overwrites
// and removes incrementally instead of being recomputed over the whole store
void launch_count_irregular(const float* rnorm, const uint32_t* meta, uint32_t r0, uint32_t n, int sign,
                            uint32_t* counter, cudaStream_t s) {
  if (!n) return;
  const uint32_t blocks = (n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592;
  count_nan_kernel<<<blocks, n < 256 ? 32 : 256, 0, s>>>(rnorm, meta, r0, n, sign, counter);
}

// bf16 shadow for the tensor pass: rows are stored NORMALISED (x / |x|) so that the
// tcgen05 contraction with a normalised query yields the cosine directly.  The shadow
// only nominates candidates; zero-norm rows become NaN and are never nominated.
__global__ void row_shadow_kernel(const float* __restrict__ E, const float* __restrict__ rnorm,
                                  __nv_bfloat16* __restrict__ E16, uint32_t ld, uint32_t ld16, uint32_t r0,
                                  uint32_t n) {
  size_t total = (size_t)n * ld16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    uint32_t r = (uint32_t)(i / ld16), c = (uint32_t)(i % ld16);
    float v = c < ld ? E[(size_t)(r0 + r) * ld + c] * rnorm[r0 + r] : 0.0f;
    E16[(size_t)(r0 + r) * ld16 + c] = __float2bfloat16_rn(v);
  }
}

void launch_prepare_rows(float* E, float* norm, float* rnorm, void* E16, uint32_t dim, uint32_t ld,
                         uint32_t ld16, uint32_t r0, uint32_t n, cudaStream_t s) {
  if (!n) return;
  row_norm_kernel<<<(n + 127) / 128, 128, 0, s>>>(E, norm, rnorm, dim, ld, r0, n);
  if (E16) {
    size_t total = (size_t)n * ld16;
    uint32_t blocks = (uint32_t)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    row_shadow_kernel<<<blocks, 256, 0, s>>>(E, rnorm, (__nv_bfloat16*)E16, ld, ld16, r0, n);
  }
}

__global__ void query_norm_kernel(const float* __restrict__ Q, float* __restrict__ qnorm,
                                  float* __restrict__ rqnorm, uint32_t nq, uint32_t qlen, uint32_t ldq) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const float* q = Q + (size_t)i * ldq;
  float acc = 0.0f;
  for (uint32_t d = 0; d < qlen; ++d) acc = ref_fold(acc, q[d], q[d]);
  float na = __fsqrt_rn(acc);
  qnorm[i] = na;
  rqnorm[i] = __frcp_rn(na);
}

void launch_prepare_queries(const float* Q, float* qnorm, float* rqnorm, uint32_t nq, uint32_t qlen,
                            uint32_t ldq, cudaStream_t s) {
  if (!nq) return;
  query_norm_kernel<<<(nq + 63) / 64, 64, 0, s>>>(Q, qnorm, rqnorm, nq, qlen, ldq);
}

// ---------------------------------------------------------------------------------
// Reference dot over min(qlen, dim) terms (Iterator::zip, index.rs:172).
__device__ __forceinline__ float ref_dot(const float* __restrict__ q, const float* __restrict__ row,
                                         uint32_t n) {
  float acc = 0.0f;
  uint32_t d = 0;
  for (; d + 4 <= n; d += 4) {
    float4 v = *reinterpret_cast<const float4*>(row + d);
    acc = ref_fold(acc, q[d], v.x);
    acc = ref_fold(acc, q[d + 1], v.y);
    acc = ref_fold(acc, q[d + 2], v.z);
    acc = ref_fold(acc, q[d + 3], v.w);
  }
  for (; d < n; ++d) acc = ref_fold(acc, q[d], row[d]);
  return acc;
}

__global__ void exact_keys_kernel_p(StoreView st, const float* __restrict__ q, const float* __restrict__ qnorm_p,
                                    uint32_t qlen, DevFilter flt, uint64_t* __restrict__ keys, uint32_t self_row,
                                    uint32_t upper_only, const uint64_t* __restrict__ row_seq, uint64_t q_seq) {
  extern __shared__ float q_s[];
  for (uint32_t d = threadIdx.x; d < qlen; d += blockDim.x) q_s[d] = q[d];
  __syncthreads();
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= st.n_rows) return;
  uint64_t key = 0;
  // pair rules of the dedup scanner (linker/dedup.rs:91-105): never the node itself, each unordered pair once
  const bool wanted = row_seq ? row_seq[row] > q_seq : (row != self_row && !(upper_only && row < self_row));
  if (wanted && row_passes(flt, st.meta, st.agent, row)) {
    uint32_t n = qlen < st.dim ? qlen : st.dim;
    float dot = ref_dot(q_s, st.E + (size_t)row * st.ld, n);
    float dist = ref_distance_from(dot, __ldg(qnorm_p), __ldg(st.norm + row));
    key = make_key(ord_from_score(ref_score_from_distance(dist)), row);
  }
  keys[row] = key;
}

void launch_exact_keys(const StoreView& st, const QueryView& qv, uint32_t q, const DevFilter& flt,
                       uint64_t* keys, cudaStream_t s, uint32_t self_row, bool upper_only, const uint64_t* row_seq,
                       uint64_t q_seq) {
  if (!st.n_rows) return;
  uint32_t threads = 128;
  exact_keys_kernel_p<<<(st.n_rows + threads - 1) / threads, threads, qv.qlen * sizeof(float), s>>>(
      st, qv.Q + (size_t)q * qv.ldq, qv.qnorm + q, qv.qlen, flt, keys, self_row, upper_only ? 1u : 0u, row_seq, q_seq);
}

size_t exact_sort_tmp_bytes(uint32_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeysDescending((void*)nullptr, bytes, (const uint64_t*)nullptr,
                                           (uint64_t*)nullptr, (int)n);
  return bytes;
}

cudaError_t exact_sort(uint64_t* keys_in, uint64_t* keys_out, uint32_t n, void* tmp, size_t tmp_bytes,
                       cudaStream_t s) {
  return cub::DeviceRadixSort::SortKeysDescending(tmp, tmp_bytes, keys_in, keys_out, (int)n, 0, 64, s);
}

__global__ void exact_count_kernel(const uint64_t* __restrict__ sorted, uint32_t n_keys, uint32_t min_ord,
                                   uint32_t* __restrict__ n_total) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_keys; i += gridDim.x * blockDim.x) {
    bool in = key_ord(sorted[i]) >= min_ord;
    bool next_in = (i + 1 < n_keys) && key_ord(sorted[i + 1]) >= min_ord;
    if (in && !next_in) *n_total = i + 1;
  }
}

__global__ void exact_emit_kernel(StoreView st, const float* __restrict__ q, const float* __restrict__ qnorm_p,
                                  uint32_t qlen, const uint64_t* __restrict__ sorted, uint32_t k,
                                  const uint32_t* __restrict__ n_total, uint32_t* __restrict__ rows,
                                  float* __restrict__ score, float* __restrict__ dist,
                                  uint8_t* __restrict__ ids, uint32_t* __restrict__ n_out) {
  uint32_t n = *n_total < k ? *n_total : k;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_out = n;
  if (i >= n) return;
  uint32_t row = key_row(sorted[i]);
  uint32_t m = qlen < st.dim ? qlen : st.dim;
  float dot = ref_dot(q, st.E + (size_t)row * st.ld, m);
  float d = ref_distance_from(dot, __ldg(qnorm_p), __ldg(st.norm + row));
  rows[i] = row;
  dist[i] = d;
  score[i] = ref_score_from_distance(d);
  if (ids) {
    const uint4* src = reinterpret_cast<const uint4*>(st.ids + (size_t)row * 16);
    *reinterpret_cast<uint4*>(ids + (size_t)i * 16) = *src;
  }
}

void launch_exact_emit(const StoreView& st, const QueryView& qv, uint32_t q, const uint64_t* sorted,
                       uint32_t n_keys, uint32_t k, uint32_t min_ord, uint32_t* rows, float* score,
                       float* dist, uint8_t* ids, uint32_t* n_out, uint32_t* n_total, cudaStream_t s) {
  cudaMemsetAsync(n_total, 0, sizeof(uint32_t), s);
  cudaMemsetAsync(n_out, 0, sizeof(uint32_t), s);
  if (!n_keys || !k) return;
  uint32_t blocks = (n_keys + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  exact_count_kernel<<<blocks, 256, 0, s>>>(sorted, n_keys, min_ord, n_total);
  uint32_t kk = k < n_keys ? k : n_keys;
  exact_emit_kernel<<<(kk + 63) / 64, 64, 0, s>>>(st, qv.Q + (size_t)q * qv.ldq, qv.qnorm + q, qv.qlen,
                                                 sorted, k, n_total, rows, score, dist, ids, n_out);
}

// ---------------------------------------------------------------------------------
__global__ void gather_rows_kernel(StoreView src, const uint64_t* __restrict__ src_seq, float* E, float* norm,
                                   float* rnorm, uint32_t* meta, uint32_t* agent, uint8_t* ids, __nv_bfloat16* E16,
                                   uint64_t* seq, const uint32_t* __restrict__ live, uint32_t n_live) {
  // one warp per destination row
  uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_live) return;
  uint32_t r = live[w];
  const float4* s4 = reinterpret_cast<const float4*>(src.E + (size_t)r * src.ld);
  float4* d4 = reinterpret_cast<float4*>(E + (size_t)w * src.ld);
  for (uint32_t i = lane; i < src.ld / 4; i += 32) d4[i] = s4[i];
  if (E16 && src.E16) {
    const uint4* s16 = reinterpret_cast<const uint4*>((const __nv_bfloat16*)src.E16 + (size_t)r * src.ld16);
    uint4* d16 = reinterpret_cast<uint4*>(E16 + (size_t)w * src.ld16);
    for (uint32_t i = lane; i < src.ld16 / 8; i += 32) d16[i] = s16[i];
  }
  if (lane == 0) {
    norm[w] = src.norm[r];
    rnorm[w] = src.rnorm[r];
    meta[w] = src.meta[r];
    agent[w] = src.agent[r];
    if (seq && src_seq) seq[w] = src_seq[r];
    *reinterpret_cast<uint4*>(ids + (size_t)w * 16) = *reinterpret_cast<const uint4*>(src.ids + (size_t)r * 16);
  }
}

void launch_gather_rows(const StoreView& src, const uint64_t* src_seq, float* E, float* norm, float* rnorm,
                        uint32_t* meta, uint32_t* agent, uint8_t* ids, void* E16, uint64_t* seq,
                        const uint32_t* live, uint32_t n_live, cudaStream_t s) {
  if (!n_live) return;
  uint32_t threads = 256, warps_per_block = threads / 32;
  gather_rows_kernel<<<(n_live + warps_per_block - 1) / warps_per_block, threads, 0, s>>>(
      src, src_seq, E, norm, rnorm, meta, agent, ids, (__nv_bfloat16*)E16, seq, live, n_live);
}

}  // namespace cx
