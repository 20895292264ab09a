// cx_kernels.h -- host-side launch interface of the scan kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>

#include "cx_common.cuh"

namespace cx {

// The opt-in dynamic shared memory limit of a kernel is a per-device function attribute shared by every host thread:
// it is only ever RAISED (under a lock), so a thread that launches with less can never find it lowered under its feet
// by a concurrent caller (or a recorded graph find it lowered at replay).  The common case is one relaxed load.
template <auto Fn>
cudaError_t raise_dynamic_smem(size_t smem, bool prefer_smem_carveout = false) {
  static std::atomic<size_t> allowed[64];
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (allowed[dev].load(std::memory_order_acquire) >= smem && smem != 0) return cudaSuccess;
  std::lock_guard<std::mutex> lk(mu);
  if (allowed[dev].load(std::memory_order_relaxed) >= smem && smem != 0) return cudaSuccess;
  e = cudaFuncSetAttribute(Fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (prefer_smem_carveout) (void)cudaFuncSetAttribute(Fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  allowed[dev].store(smem, std::memory_order_release);
  return cudaSuccess;
}

// Device view of the embedding store.
struct StoreView {
  const float* E;          // [n_rows][ld] fp32, rows 16 B aligned, zero padded to ld
  const float* norm;       // [n_rows] sqrt(sum x^2) in reference order (index.rs:174)
  const float* rnorm;      // [n_rows] 1/norm (fast passes only)
  const uint32_t* meta;    // [n_rows] META_* word
  const uint32_t* agent;   // [n_rows] interned agent id
  const uint8_t* ids;      // [n_rows][16]
  const void* E16;         // [n_rows][ld16] bf16 shadow (tensor pass) or nullptr
  uint32_t n_rows, dim, ld, ld16;
};

// Queries prepared on device for one call.
struct QueryView {
  const float* Q;     // [nq][ldq] fp32 zero padded (ldq >= max(ld, qlen rounded to 4))
  const float* qnorm; // [nq] reference-order norm over qlen elements
  const float* rqnorm;// [nq] 1/qnorm
  uint32_t nq, qlen, ldq;
};

// Candidates nominated by a fast pass: one merged, unordered key list per query.
// cnt and gtau must be zero when a pass starts; the select kernel re-zeroes them.
struct CandView {
  uint64_t* keys;   // [nq][cap]; the list of query q starts at keys + (q - q_base) * cap
  uint32_t q_base = 0;
  uint32_t* cnt;    // [nq] entries appended (may exceed cap: the excess was dropped)
  uint64_t* gtau;   // [nq] running cut-off shared by all producer groups; at the end of the
                    //      pass it bounds (as a key) every row that is NOT in the list
  uint32_t cap;     // list capacity per query
  uint32_t G, KP;   // producer groups, keys kept per group
};

struct ResultView {
  uint32_t* rows;   // [nq][k]
  float* score;     // [nq][k]
  float* dist;      // [nq][k]
  uint8_t* ids;     // [nq][k][16]
  uint32_t* n;      // [nq]
  uint32_t* ok;     // [nq] 1 = verified exact, 0 = caller must fall back
  uint32_t k;
};

// K4: norms (reference order), reciprocal norms, bf16 shadow rows for rows [r0, r0+n)
void launch_prepare_rows(float* E, float* norm, float* rnorm, void* E16, uint32_t dim, uint32_t ld,
                         uint32_t ld16, uint32_t r0, uint32_t n, cudaStream_t s);
// *counter += sign * (rows of [r0, r0+n) with a NaN reciprocal norm that are not removed)
void launch_count_irregular(const float* rnorm, const uint32_t* meta, uint32_t r0, uint32_t n, int sign,
                            uint32_t* counter, cudaStream_t s);
// queries: reference-order norm over qlen, reciprocal
void launch_prepare_queries(const float* Q, float* qnorm, float* rqnorm, uint32_t nq, uint32_t qlen,
                            uint32_t ldq, cudaStream_t s);

// K1: streaming fp32 pass.  Handles queries [q0, q0+nq_pass) with nq_pass <= 8.
// Returns the number of producer groups it will write (G) for a store of n_rows.
uint32_t stream_scan_groups(uint32_t n_rows, int sm_count);
size_t stream_scan_smem(uint32_t ld, uint32_t nq_pass, uint32_t KP);
size_t stream_scan_half_smem(uint32_t ld16, uint32_t nq_pass, uint32_t KP);  // bf16-shadow variant (nq_pass <= 4)
// qmap (device, optional, q0 must be 0): pass-local query b is query qmap[b]; its candidate list is slot b
// thr_cos (optional): threshold mode -- every row whose approximate cosine reaches *thr_cos is nominated
// (no top-KP cut-off; qv.qnorm must already hold the query norms); cv.gtau is not used.
cudaError_t launch_stream_scan(const StoreView& st, const QueryView& qv, uint32_t q0, uint32_t nq_pass,
                               const DevFilter& flt, const CandView& cv, int sm_count, cudaStream_t s,
                               const uint32_t* qmap = nullptr, const float* thr_cos = nullptr, bool half = false);
// half: stream the normalised bf16 shadow (st.E16) instead of the fp32 rows; nq_pass <= 4, no threshold mode

// K5 + K3: merge candidate lists, exact rescore, verify, emit results.
// eps_cos: bound on |approx - reference| cosine for the pass that produced the candidates.
// It also computes the query norms (qv.qnorm / qv.rqnorm are written) and re-zeroes cv.cnt / cv.gtau.
// scale_by_rqn: the pass's keys are cosine * |q| (streaming pass) instead of cosine.
size_t select_smem(uint32_t cap, uint32_t ld);
// diagnostics: queries that failed verification on the current device since process start, by reason
// ([0] list overflow, [1] near-ties denser than the rescored band, [2] other)
void select_why_read(uint64_t out[3]);
cudaError_t launch_select_rescore(const StoreView& st, const QueryView& qv, uint32_t q0, uint32_t nq,
                                  const CandView& cv, const ResultView& rv, float eps_cos, int scale_by_rqn,
                                  cudaStream_t s, const uint32_t* qmap = nullptr);

// Threshold scans: rescore ALL nominees of each query exactly, keep `score >= threshold`
// (index.rs:385), order by (score desc, row asc), write up to rv.k of them; total[q] = how many
// qualify; rv.ok[q] = 0 if the nominee list overflowed or more than 2048 rows qualify.
// qv.qnorm must hold the query norms.  self_rows (optional): skip that row; upper_only: keep
// only rows above it (the dedup scanner's unordered pairs, linker/dedup.rs:96-105).
size_t threshold_rescore_smem(uint32_t ld);
// q_seq / row_seq (multi-device pair scans): keep row r only if row_seq[r] > q_seq[q].
cudaError_t launch_threshold_rescore(const StoreView& st, const QueryView& qv, uint32_t q0, uint32_t nq,
                                     const CandView& cv, const ResultView& rv, uint32_t* total, float threshold,
                                     const uint32_t* self_rows, bool upper_only, cudaStream_t s,
                                     const uint64_t* q_seq = nullptr, const uint64_t* row_seq = nullptr);

// K2: tcgen05 bf16 pass over the normalised shadow matrix.  Q16 = normalised bf16 queries
// [round_up(nq_total,128)][ld16] made by launch_query_bf16.  One launch serves at most
// sm_count*128 queries starting at q0; `lists` is scratch of tensor_scratch_bytes().
// cv.KP must be tensor_keep(k) (32 / 64 / 128 scores tracked per query in registers).
uint32_t tensor_keep(uint32_t k);
bool tensor_scan_eligible(uint32_t ld16, uint32_t k);
// Per-index tuning of the tensor pass, read-only while searches run: pair = CTA-pair (cta_group::2) form,
// epi_warps = 8 | 16 epilogue warps per CTA.  debug is honoured only by -DCX_PROBE builds (measurement
// hook, results become wrong: 1 no epilogue work, 2 no hit handling, 4 no E traffic).
struct TensorTuning {
  int pair = 0;
  int epi_warps = 8;
  int debug = 0;
  int use_leftover_sms = 1;  // give the SMs the (query tiles x row splits) grid leaves over a share of the rows
};
void tensor_report_clock();  // CX_PROBE builds: print the effective SM clock of block 0 in the last debug-mode launch
size_t tensor_scratch_bytes(int sm_count);
void launch_query_bf16(const float* Q, uint32_t ldq, uint32_t dim, uint32_t nq, uint32_t nq_pad, void* Q16,
                       uint32_t ld16, cudaStream_t s);
// Bootstrap of the per-query cut-off: scores of `n_slots` sampled row tiles are dumped
// ([nq][n_slots*256] floats) and cv.gtau[q] is set to the cv.KP-th best of them.
uint32_t tensor_sample_tiles(uint32_t n_rows, uint32_t nq);
cudaError_t launch_tensor_bootstrap(const StoreView& st, const void* Q16, uint32_t q0, uint32_t nq,
                                    const DevFilter& flt, bool check_rows, const CandView& cv, float* dump,
                                    uint32_t n_slots, int sm_count, cudaStream_t s, const TensorTuning& tune);
// One phase of the scan: row tiles [tile0, tile0 + n_tiles) of tensor_tiles(n_rows) (256 rows each).
// check_rows: a filter is active or rows were removed -> test metadata before nominating a row
cudaError_t launch_tensor_scan(const StoreView& st, const void* Q16, uint32_t q0, uint32_t nq,
                               const DevFilter& flt, bool check_rows, const CandView& cv, uint64_t* lists,
                               uint32_t tile0, uint32_t n_tiles, int sm_count, cudaStream_t s,
                               const TensorTuning& tune, bool static_tau = false);
// Threshold scans: fix every query's cut-off at the approximate cosine thr_cos (then scan with static_tau).
void launch_fill_tau(const CandView& cv, uint32_t q0, uint32_t nq, float thr_cos, cudaStream_t s);
uint32_t tensor_tiles(uint32_t n_rows);
// Between phases: raise each query's cut-off to the cv.KP-th best score nominated so far minus `margin`.
cudaError_t launch_tau_refine(const CandView& cv, uint32_t q0, uint32_t nq, float margin, cudaStream_t s);

// Exact path: every row scored with reference arithmetic -> keys[n_rows] for one query
// self_row / upper_only: the dedup scanner's pair rules (skip the query's own row; keep only rows above it)
// row_seq / q_seq (multi-device pair scans): keep row r only if row_seq[r] > q_seq
void launch_exact_keys(const StoreView& st, const QueryView& qv, uint32_t q, const DevFilter& flt,
                       uint64_t* keys, cudaStream_t s, uint32_t self_row = 0xFFFFFFFFu, bool upper_only = false,
                       const uint64_t* row_seq = nullptr, uint64_t q_seq = 0);
// sort keys descending (cub radix sort); tmp sized by exact_sort_tmp_bytes
size_t exact_sort_tmp_bytes(uint32_t n);
cudaError_t exact_sort(uint64_t* keys_in, uint64_t* keys_out, uint32_t n, void* tmp, size_t tmp_bytes,
                       cudaStream_t s);
// emit the first min(k, #keys >= min_ord) results of a sorted key array for query q
void launch_exact_emit(const StoreView& st, const QueryView& qv, uint32_t q, const uint64_t* sorted,
                       uint32_t n_keys, uint32_t k, uint32_t min_ord, uint32_t* rows, float* score,
                       float* dist, uint8_t* ids, uint32_t* n_out, uint32_t* n_total, cudaStream_t s);

// auto-link candidate post-pass over device-resident search results (cx_merge.cu)
void launch_autolink_filter(const uint32_t* rows, const float* score, const uint32_t* n, const uint32_t* self_rows,
                            const uint8_t* ids, uint32_t B, uint32_t k, float threshold, uint32_t max_edges,
                            uint32_t* out_rows, float* out_score, uint8_t* out_ids, uint32_t* out_n,
                            cudaStream_t s);

// dst[i] = base + i
void launch_iota(uint32_t* dst, uint32_t n, uint32_t base, cudaStream_t s);
// Dedup scan: per-node partner lists [B][kd] -> dense (a row, b row, score bits) triples.  Nodes whose own
// row (r0 + i) was removed contribute nothing (n and tot are zeroed in place).  off [B+1] receives the
// exclusive prefix sum of the counts (off[B] = pairs in this block); at most `limit` triples are written.
void launch_compact_pairs(const uint32_t* rows, const float* score, uint32_t* n, uint32_t* tot, const uint32_t* meta,
                          uint32_t r0, uint32_t B, uint32_t kd, uint32_t* off, uint32_t* pairs, uint32_t limit,
                          cudaStream_t s);

void launch_compact_offsets(uint32_t* n, const uint32_t* dead, uint32_t B, uint32_t* off, cudaStream_t s);

// compaction helper for rebuild(): dst[i] = src[live[i]] for all per-row arrays
void launch_gather_rows(const StoreView& src, const uint64_t* src_seq, float* E, float* norm, float* rnorm,
                        uint32_t* meta, uint32_t* agent, uint8_t* ids, void* E16, uint64_t* seq,
                        const uint32_t* live, uint32_t n_live, cudaStream_t s);

}  // namespace cx
