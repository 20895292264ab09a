// cx_extract.cu -- the two callers' data formats either side of the scan:
//
//  * Node extractor + bulk load.  At start-up the reference decodes every node of the redb `nodes`
//    table (bincode 1.3, fixint, little endian) into a full `Node`, sorts them newest first and
//    inserts the embeddings one by one (serve.rs:105-123, api.rs:55-69,
//    storage/redb_storage.rs:670-734).  Here the raw table values go to the device as one blob; one
//    thread per node walks the variable-length fields up to the `Option<Vec<f32>>` (layout pinned by
//    the reference's golden bytes, storage/redb_storage.rs:1834-1856), a second kernel gathers the
//    embeddings -- in the reference's insertion order -- straight into the store.  The floats never
//    visit the host.  Nodes the walk cannot decide (non-empty `metadata`: bincode does not describe
//    serde_json::Value; unusual timestamps) are flagged for the caller's own decoder.
//
//  * Query-time score decay (vector/scoring.rs:84-114) over the candidates of a search, with the
//    re-ranking of http/routes.rs:945-949, as one kernel.
#include <algorithm>
#include <memory>
#include <numeric>

#include "cx_index.h"

namespace cx {

// ---- bincode walk ------------------------------------------------------------------------------
struct NodeRecord {
  int64_t created_ns;        // created_at, nanoseconds since the epoch
  int64_t last_accessed_ns;  // last_accessed_at
  uint64_t access_count;
  uint64_t emb_off;          // byte offset of the first float inside the blob
  uint32_t emb_len;          // floats
  uint32_t status;           // CX_NODE_*
  uint8_t id[16];
};

struct Cursor {
  const uint8_t* p;
  uint64_t off, end;
  bool bad;
  __device__ uint8_t u8() {
    if (off + 1 > end) {
      bad = true;
      return 0;
    }
    return p[off++];
  }
  __device__ uint64_t u64() {
    if (off + 8 > end) {
      bad = true;
      return 0;
    }
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v |= (uint64_t)p[off + i] << (8 * i);
    off += 8;
    return v;
  }
  __device__ void skip(uint64_t n) {
    if (n > end - off) bad = true;
    else off += n;
  }
};

// Rust's String must be UTF-8: serde rejects anything else and the reference then skips the record
// (storage/redb_storage.rs:707-710).  Standard DFA-free check: lengths, continuation bytes, overlongs,
// surrogates, > U+10FFFF.
__device__ bool utf8_ok(const uint8_t* s, uint64_t n) {
  uint64_t i = 0;
  while (i < n) {
    const uint8_t c = s[i];
    if (c < 0x80) {
      ++i;
      continue;
    }
    uint32_t need, min_cp, cp;
    if ((c & 0xE0) == 0xC0) need = 1, min_cp = 0x80, cp = c & 0x1F;
    else if ((c & 0xF0) == 0xE0) need = 2, min_cp = 0x800, cp = c & 0x0F;
    else if ((c & 0xF8) == 0xF0) need = 3, min_cp = 0x10000, cp = c & 0x07;
    else return false;
    for (uint32_t k = 1; k <= need; ++k) {
      if (i + k >= n) return false;
      const uint8_t x = s[i + k];
      if ((x & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (x & 0x3F);
    }
    if (cp < min_cp || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
    i += need + 1;
  }
  return true;
}

// String: u64 length + bytes
__device__ bool walk_string(Cursor& c, const uint8_t** s = nullptr, uint64_t* n = nullptr) {
  const uint64_t len = c.u64();
  if (c.bad) return false;
  const uint64_t at = c.off;
  c.skip(len);
  if (c.bad) return false;
  if (!utf8_ok(c.p + at, len)) {
    c.bad = true;
    return false;
  }
  if (s) *s = c.p + at;
  if (n) *n = len;
  return true;
}

__device__ bool walk_opt_string(Cursor& c) {
  const uint8_t tag = c.u8();
  if (c.bad || tag > 1) {
    c.bad = true;
    return false;
  }
  return tag ? walk_string(c) : true;
}

__device__ int64_t days_from_civil(int64_t y, int64_t m, int64_t d) {
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const int64_t yoe = y - era * 400;
  const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}

// chrono's serde form of DateTime<Utc>: RFC 3339, "YYYY-MM-DDTHH:MM:SS[.fraction](Z|+hh:mm|-hh:mm)".
// Returns false for anything else (the node is then handed to the caller's decoder, never guessed).
__device__ bool parse_rfc3339(const uint8_t* s, uint64_t n, int64_t* out_ns) {
  auto dig = [&](uint64_t i) { return i < n && s[i] >= '0' && s[i] <= '9'; };
  auto num = [&](uint64_t i, int k, int64_t* v) {
    int64_t x = 0;
    for (int j = 0; j < k; ++j) {
      if (!dig(i + j)) return false;
      x = x * 10 + (s[i + j] - '0');
    }
    *v = x;
    return true;
  };
  int64_t Y, M, D, h, m, sec;
  if (n < 20) return false;
  if (!num(0, 4, &Y) || s[4] != '-' || !num(5, 2, &M) || s[7] != '-' || !num(8, 2, &D)) return false;
  if (s[10] != 'T' || !num(11, 2, &h) || s[13] != ':' || !num(14, 2, &m) || s[16] != ':' || !num(17, 2, &sec)) return false;
  if (M < 1 || M > 12 || D < 1 || D > 31 || h > 23 || m > 59 || sec > 59) return false;  // leap seconds: host decode
  uint64_t i = 19;
  int64_t frac = 0;
  if (i < n && s[i] == '.') {
    ++i;
    int nd = 0;
    while (dig(i)) {
      if (nd < 9) frac = frac * 10 + (s[i] - '0');
      ++nd;
      ++i;
    }
    if (nd == 0 || nd > 9) return false;
    for (; nd < 9; ++nd) frac *= 10;
  }
  int64_t off_s = 0;
  if (i < n && s[i] == 'Z') {
    ++i;
  } else if (i < n && (s[i] == '+' || s[i] == '-')) {
    const int64_t sign = s[i] == '-' ? -1 : 1;
    int64_t oh, om;
    if (!num(i + 1, 2, &oh) || i + 3 >= n || s[i + 3] != ':' || !num(i + 4, 2, &om) || oh > 23 || om > 59) return false;
    off_s = sign * (oh * 3600 + om * 60);
    i += 6;
  } else {
    return false;
  }
  if (i != n) return false;
  const int64_t secs = days_from_civil(Y, M, D) * 86400 + h * 3600 + m * 60 + sec - off_s;
  if (secs > 9000000000ll || secs < -9000000000ll) return false;  // keeps the nanosecond count inside int64
  *out_ns = secs * 1000000000ll + frac;
  return true;
}

// Node (types.rs:26-68): id, kind, data{title, body, metadata, tags}, embedding, source{agent, session,
// channel}, importance, access_count, last_accessed_at, created_at, updated_at, deleted.
__global__ void node_walk_kernel(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ offsets, uint64_t n,
                                 uint32_t dim, NodeRecord* __restrict__ rec) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  NodeRecord r;
  memset(&r, 0, sizeof r);
  Cursor c{blob, offsets[i], offsets[i + 1], false};
  auto finish = [&](uint32_t status) {
    r.status = status;
    rec[i] = r;
  };
  if (c.end < c.off) return finish(CX_NODE_CORRUPT);
  // id: uuid as a byte sequence -- u64 length (16) + bytes
  if (c.u64() != 16 || c.bad) return finish(CX_NODE_CORRUPT);
  for (int b = 0; b < 16; ++b) r.id[b] = c.u8();
  if (c.bad) return finish(CX_NODE_CORRUPT);
  if (!walk_string(c)) return finish(CX_NODE_CORRUPT);  // kind
  if (!walk_string(c)) return finish(CX_NODE_CORRUPT);  // title
  if (!walk_string(c)) return finish(CX_NODE_CORRUPT);  // body
  const uint64_t n_meta = c.u64();
  if (c.bad) return finish(CX_NODE_CORRUPT);
  // HashMap<String, serde_json::Value>: bincode writes a Value without a type tag, so a non-empty map cannot
  // be skipped from the bytes alone
  if (n_meta != 0) return finish(CX_NODE_NEEDS_HOST_DECODE);
  const uint64_t n_tags = c.u64();
  if (c.bad || n_tags > (c.end - c.off) / 8) return finish(CX_NODE_CORRUPT);
  for (uint64_t t = 0; t < n_tags; ++t)
    if (!walk_string(c)) return finish(CX_NODE_CORRUPT);
  const uint8_t has_emb = c.u8();
  if (c.bad || has_emb > 1) return finish(CX_NODE_CORRUPT);
  uint64_t emb_len = 0;
  if (has_emb) {
    emb_len = c.u64();
    if (c.bad || emb_len > (c.end - c.off) / 4) return finish(CX_NODE_CORRUPT);
    r.emb_off = c.off;
    r.emb_len = (uint32_t)(emb_len > 0xFFFFFFFFull ? 0xFFFFFFFFull : emb_len);
    c.skip(emb_len * 4);
  }
  if (!walk_string(c)) return finish(CX_NODE_CORRUPT);      // source.agent
  if (!walk_opt_string(c)) return finish(CX_NODE_CORRUPT);  // source.session
  if (!walk_opt_string(c)) return finish(CX_NODE_CORRUPT);  // source.channel
  c.skip(4);                                                // importance f32
  r.access_count = c.u64();
  if (c.bad) return finish(CX_NODE_CORRUPT);
  const uint8_t* ts[3];
  uint64_t tn[3];
  for (int t = 0; t < 3; ++t)  // last_accessed_at, created_at, updated_at
    if (!walk_string(c, &ts[t], &tn[t])) return finish(CX_NODE_CORRUPT);
  const uint8_t deleted = c.u8();
  if (c.bad || deleted > 1) return finish(CX_NODE_CORRUPT);
  int64_t upd;
  if (!parse_rfc3339(ts[0], tn[0], &r.last_accessed_ns) || !parse_rfc3339(ts[1], tn[1], &r.created_ns) ||
      !parse_rfc3339(ts[2], tn[2], &upd))
    return finish(CX_NODE_NEEDS_HOST_DECODE);
  if (deleted) return finish(CX_NODE_DELETED);  // list_nodes(NodeFilter::new()) leaves tombstones out (redb_storage.rs:347)
  if (!has_emb) return finish(CX_NODE_NO_EMBEDDING);
  if (emb_len != dim) return finish(CX_NODE_DIM_MISMATCH);  // index.insert(..) is Err and the loop moves on (serve.rs:112-114)
  finish(CX_NODE_OK);
}

// rows[j][0..dim) = the embedding of node order[j]; the floats sit at arbitrary byte offsets in the blob
__global__ void node_gather_kernel(const uint8_t* __restrict__ blob, const NodeRecord* __restrict__ rec,
                                   const uint32_t* __restrict__ order, uint64_t m, uint32_t dim,
                                   float* __restrict__ rows, uint8_t* __restrict__ ids) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (w >= m) return;
  const NodeRecord& r = rec[order[w]];
  const uint8_t* src = blob + r.emb_off;
  const uint32_t mis = (uint32_t)((uintptr_t)src & 3u);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(src - mis);  // aligned words; funnel-shift the pieces together
  float* dst = rows + w * dim;
  for (uint32_t d = lane; d < dim; d += 32) {
    const uint32_t lo = base[d];
    const uint32_t v = mis ? __funnelshift_r(lo, base[d + 1], 8 * mis) : lo;
    dst[d] = __uint_as_float(v);
  }
  if (ids && lane < 16) ids[w * 16 + lane] = r.id[lane];
}

// ---- score decay -----------------------------------------------------------------------------------
// vector/scoring.rs:84-114, operation for operation: days idle and the two factors in f64, the blend in f32.
__device__ __forceinline__ float decay_one(float raw, int64_t idle_seconds, uint64_t access_count, double kind_rate,
                                           const cx_decay_config& cfg, float bias) {
  if (!cfg.enabled || bias == 0.0f) return raw;
  const double days_idle = (double)(idle_seconds > 0 ? idle_seconds : 0) / 86400.0;
  const double effective_days = days_idle < cfg.max_age_days ? days_idle : cfg.max_age_days;  // f64::min
  double t = exp(-kind_rate * effective_days);
  t = t > cfg.min_factor ? t : cfg.min_factor;  // f64::max
  const float temporal = (float)t;
  double e = 1.0 + (double)access_count * cfg.echo_weight;
  e = e < cfg.echo_cap ? e : cfg.echo_cap;
  const float echo = (float)e;
  const float keep = __fmul_rn(raw, __fsub_rn(1.0f, bias));
  const float moved = __fmul_rn(__fmul_rn(__fmul_rn(raw, temporal), echo), bias);
  return __fadd_rn(keep, moved);
}

// one CTA per segment (the candidates of one query): decay, then the stable descending re-rank of
// routes.rs:945-949 by rank counting (NaN compares Equal there; here NaN sorts last, like the index)
__global__ void score_decay_kernel(cx_decay_config cfg, float bias, uint64_t n, uint32_t seg_len,
                                   const float* __restrict__ raw, const int64_t* __restrict__ idle,
                                   const uint64_t* __restrict__ acc, const double* __restrict__ rate,
                                   float* __restrict__ out, uint32_t* __restrict__ order) {
  extern __shared__ float s_sc[];
  const uint64_t base = (uint64_t)blockIdx.x * seg_len;
  const uint32_t len = (uint32_t)(base + seg_len <= n ? seg_len : n - base);
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
    const float v = decay_one(raw[base + i], idle[base + i], acc[base + i], rate[base + i], cfg, bias);
    out[base + i] = v;
    if (order) s_sc[i] = v;
  }
  if (!order) return;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
    const uint32_t mine = ord_from_score(s_sc[i]);
    uint32_t rank = 0;
    for (uint32_t j = 0; j < len; ++j) {
      const uint32_t o = ord_from_score(s_sc[j]);
      rank += (o > mine) || (o == mine && j < i);
    }
    order[base + rank] = i;
  }
}

}  // namespace cx

using namespace cx;

namespace {

struct DeviceBlock {
  void* p = nullptr;
  ~DeviceBlock() {
    if (p) cudaFree(p);
  }
};

// walk all nodes; rec (host, n entries) receives the records; d_blob / d_rec stay on the device for the gather
cx_status walk_nodes(const uint8_t* values, const uint64_t* offsets, uint64_t n, uint32_t dim, cudaStream_t s,
                     DeviceBlock& d_blob, DeviceBlock& d_off, DeviceBlock& d_rec, std::vector<NodeRecord>& rec) {
  const uint64_t total = offsets[n];
  for (uint64_t i = 0; i < n; ++i)
    if (offsets[i + 1] < offsets[i]) return fail(CX_ERR_VALIDATION, "offsets must be non-decreasing");
  CU(cudaMalloc(&d_blob.p, total + 8));  // + slack: the gather reads whole aligned words
  CU(cudaMalloc(&d_off.p, (n + 1) * 8));
  CU(cudaMalloc(&d_rec.p, n * sizeof(NodeRecord)));
  CU(cudaMemsetAsync((uint8_t*)d_blob.p + total, 0, 8, s));
  CU(cudaMemcpyAsync(d_blob.p, values, total, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d_off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, s));
  node_walk_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>((const uint8_t*)d_blob.p, (const uint64_t*)d_off.p, n, dim,
                                                              (NodeRecord*)d_rec.p);
  CU(cudaGetLastError());
  rec.resize(n);
  CU(cudaMemcpyAsync(rec.data(), d_rec.p, n * sizeof(NodeRecord), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return CX_OK;
}

}  // namespace

// Standalone extraction to host arrays (what the start-up loop needs from every node, without building a Node).
extern "C" cx_status cx_extract_embeddings(const uint8_t* values, const uint64_t* offsets, uint64_t n, uint32_t dim,
                                           int device, uint8_t* out_ids, float* out_rows, int64_t* out_created_ns,
                                           int64_t* out_last_accessed_ns, uint64_t* out_access_count,
                                           uint8_t* out_status) {
  cx::CallerDevice keep_callers_device;
  if (!n) return CX_OK;
  if (!values || !offsets || !out_status) return fail(CX_ERR_VALIDATION, "null argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(CX_ERR_CUDA, "no CUDA device available; cortex_gpu has no CPU fallback");
  if (device < 0 || device >= n_dev) return fail(CX_ERR_VALIDATION, "device %d out of range", device);
  CU(cudaSetDevice(device));
  cudaStream_t s;
  CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  std::unique_ptr<CUstream_st, void (*)(cudaStream_t)> sg(s, [](cudaStream_t x) { cudaStreamDestroy(x); });
  DeviceBlock d_blob, d_off, d_rec, d_rows, d_order;
  std::vector<NodeRecord> rec;
  cx_status st = walk_nodes(values, offsets, n, dim, s, d_blob, d_off, d_rec, rec);
  if (st != CX_OK) return st;
  std::vector<uint32_t> order;
  for (uint64_t i = 0; i < n; ++i) {
    out_status[i] = (uint8_t)rec[i].status;
    if (out_ids) memcpy(out_ids + 16 * i, rec[i].id, 16);
    if (out_created_ns) out_created_ns[i] = rec[i].created_ns;
    if (out_last_accessed_ns) out_last_accessed_ns[i] = rec[i].last_accessed_ns;
    if (out_access_count) out_access_count[i] = rec[i].access_count;
    if (rec[i].status == CX_NODE_OK) order.push_back((uint32_t)i);
  }
  if (out_rows && !order.empty()) {
    // embeddings of the usable nodes, gathered densely, then scattered to the caller's [n][dim] rows
    const uint64_t m = order.size();
    CU(cudaMalloc(&d_rows.p, m * dim * 4));
    CU(cudaMalloc(&d_order.p, m * 4));
    CU(cudaMemcpyAsync(d_order.p, order.data(), m * 4, cudaMemcpyHostToDevice, s));
    node_gather_kernel<<<(unsigned)((m + 7) / 8), 256, 0, s>>>((const uint8_t*)d_blob.p, (const NodeRecord*)d_rec.p,
                                                              (const uint32_t*)d_order.p, m, dim, (float*)d_rows.p, nullptr);
    CU(cudaGetLastError());
    std::vector<float> tmp(m * dim);
    CU(cudaMemcpyAsync(tmp.data(), d_rows.p, m * dim * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (uint64_t j = 0; j < m; ++j) memcpy(out_rows + (size_t)order[j] * dim, tmp.data() + j * dim, (size_t)dim * 4);
  }
  return CX_OK;
}

// The start-up loop (serve.rs:105-123 / api.rs:55-69) in one call: every live node that carries an embedding
// of the index's dimension is inserted, newest first (list_nodes sorts by created_at descending, stable:
// redb_storage.rs:728); the embeddings go from the uploaded blob into the store on the device.
extern "C" cx_status cx_load_nodes(cx_index* h, const uint8_t* values, const uint64_t* offsets, uint64_t n,
                                   uint8_t* out_status, uint64_t out_counts[6]) {
  cx::CallerDevice keep_callers_device;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (out_counts) memset(out_counts, 0, 6 * sizeof(uint64_t));
  if (!n) return CX_OK;
  if (!values || !offsets) return fail(CX_ERR_VALIDATION, "null argument");
  const int device = h->device;  // a multi-device index stages on devices[0]
  CU(cudaSetDevice(device));
  cudaStream_t s;
  CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  std::unique_ptr<CUstream_st, void (*)(cudaStream_t)> sg(s, [](cudaStream_t x) { cudaStreamDestroy(x); });
  DeviceBlock d_blob, d_off, d_rec, d_rows, d_order;
  std::vector<NodeRecord> rec;
  cx_status st = walk_nodes(values, offsets, n, h->dim, s, d_blob, d_off, d_rec, rec);
  if (st != CX_OK) return st;
  std::vector<uint32_t> order;
  for (uint64_t i = 0; i < n; ++i) {
    if (out_status) out_status[i] = (uint8_t)rec[i].status;
    if (out_counts && rec[i].status < 6) out_counts[rec[i].status]++;
    if (rec[i].status == CX_NODE_OK) order.push_back((uint32_t)i);
  }
  if (order.empty()) return CX_OK;
  std::stable_sort(order.begin(), order.end(),
                   [&](uint32_t a, uint32_t b) { return rec[a].created_ns > rec[b].created_ns; });
  const uint64_t m = order.size();
  std::vector<uint8_t> ids(m * 16);
  for (uint64_t j = 0; j < m; ++j) memcpy(&ids[16 * j], rec[order[j]].id, 16);
  CU(cudaMalloc(&d_rows.p, m * h->dim * 4));
  CU(cudaMalloc(&d_order.p, m * 4));
  CU(cudaMemcpyAsync(d_order.p, order.data(), m * 4, cudaMemcpyHostToDevice, s));
  node_gather_kernel<<<(unsigned)((m + 7) / 8), 256, 0, s>>>((const uint8_t*)d_blob.p, (const NodeRecord*)d_rec.p,
                                                            (const uint32_t*)d_order.p, m, h->dim, (float*)d_rows.p, nullptr);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s));
  return cx_insert_batch_device(h, ids.data(), (const float*)d_rows.p, m, h->dim);
}

// apply_score_decay (vector/scoring.rs:84-114) for n candidates, `seg_len` per query; out_order (optional):
// per segment, the candidates' positions re-ranked by the decayed score (routes.rs:945-949).
extern "C" cx_status cx_apply_score_decay(cx_index* h, const cx_decay_config* cfg, float recency_bias, uint64_t n,
                                          uint32_t seg_len, const float* raw_score, const int64_t* idle_seconds,
                                          const uint64_t* access_count, const double* kind_rate, float* out_score,
                                          uint32_t* out_order) {
  cx::CallerDevice keep_callers_device;
  if (!h || !cfg) return fail(CX_ERR_VALIDATION, "null argument");
  if (!n) return CX_OK;
  if (!raw_score || !idle_seconds || !access_count || !kind_rate || !out_score)
    return fail(CX_ERR_VALIDATION, "null argument");
  if (!seg_len) seg_len = (uint32_t)std::min<uint64_t>(n, 1024);
  if (out_order && seg_len > 8192) return fail(CX_ERR_VALIDATION, "re-ranking handles at most 8192 candidates per query");
  cx_index* c = h->shards ? nullptr : h;
  CU(cudaSetDevice(h->device));
  DeviceBlock blk;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
  const size_t o_raw = take(n * 4), o_idle = take(n * 8), o_acc = take(n * 8), o_rate = take(n * 8), o_out = take(n * 4),
               o_ord = take(out_order ? n * 4 : 0);
  // the inputs travel through a leased workspace of the index (no allocation per call once warm)
  std::unique_ptr<WsLease> lease;
  char* d = nullptr;
  cudaStream_t s = nullptr;
  std::unique_ptr<CUstream_st, void (*)(cudaStream_t)> sg(nullptr, [](cudaStream_t x) { if (x) cudaStreamDestroy(x); });
  if (c) {
    lease.reset(new WsLease(c));
    CU(lease->init());
    CU(lease->ws->ensure_aux(off, 256));
    d = (char*)lease->ws->aux;
    s = lease->ws->stream;
  } else {
    CU(cudaMalloc(&blk.p, off));
    d = (char*)blk.p;
    CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    sg.reset(s);
  }
  CU(cudaMemcpyAsync(d + o_raw, raw_score, n * 4, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d + o_idle, idle_seconds, n * 8, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d + o_acc, access_count, n * 8, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d + o_rate, kind_rate, n * 8, cudaMemcpyHostToDevice, s));
  const uint64_t n_seg = (n + seg_len - 1) / seg_len;
  score_decay_kernel<<<(unsigned)n_seg, 256, out_order ? seg_len * 4 : 0, s>>>(
      *cfg, recency_bias, n, seg_len, (const float*)(d + o_raw), (const int64_t*)(d + o_idle), (const uint64_t*)(d + o_acc),
      (const double*)(d + o_rate), (float*)(d + o_out), out_order ? (uint32_t*)(d + o_ord) : nullptr);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out_score, d + o_out, n * 4, cudaMemcpyDeviceToHost, s));
  if (out_order) CU(cudaMemcpyAsync(out_order, d + o_ord, n * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return CX_OK;
}
