// cx_search.cu -- search entry points of the C ABI: VectorIndex::search /
// search_threshold / search_batch (vector/index.rs:325-410) and the device-resident
// batch form.  Path selection per call:
//
//   small query groups  -> K1 streaming fp32 pass (cx_stream.cu)   } nominate candidates,
//   large query groups  -> K2 tcgen05 bf16 pass  (cx_tensor.cu)    } then K5/K3 select +
//                                                                  } exact rescore + verify
//   tiny indexes, huge k, threshold scans, mismatched query length, and every query
//   whose fast result could not be verified -> exact path (cx_exact.cu): reference
//   arithmetic for every row + radix sort.
//
// Whatever the path, emitted scores/distances come from the reference's operation
// sequence, and ids follow (score desc, NaN last, row asc).
#include <algorithm>
#include <memory>
#include <vector>

#include "cx_index.h"

using namespace cx;

namespace {

struct FilterHost {
  DevFilter dev;
  std::vector<uint32_t> excl_rows;
};

void build_filter(const cx_index* h, const cx_filter* f, FilterHost* out) {
  DevFilter& d = out->dev;
  memset(&d, 0, sizeof d);
  d.agent = AGENT_NONE;
  if (!f) return;
  if (f->has_kinds) {
    d.has_kinds = 1;
    for (uint32_t i = 0; i < f->n_kinds; ++i) {
      uint32_t id;
      if (f->kinds && f->kinds[i] && h->kinds.find(f->kinds[i], &id)) d.kind_mask[id >> 6] |= 1ull << (id & 63);
    }
  }
  if (f->has_source_agent) {
    d.has_agent = 1;
    uint32_t id;
    if (f->source_agent && h->agents.find(f->source_agent, &id)) d.agent = id;
  }
  if (f->has_exclude && f->exclude_ids) {
    for (uint32_t i = 0; i < f->n_exclude; ++i) {
      if (const uint32_t* row = h->id2row.find(load_id(f->exclude_ids + 16 * i))) out->excl_rows.push_back(*row);
    }
  }
}

// keys kept per producer group: k plus a margin for near-ties, bounded so that the
// merged list of all groups still fits the select kernel's shared memory
uint32_t keep_count(uint32_t k, uint32_t G) {
  uint32_t margin = k / 4;
  if (margin < 16) margin = 16;
  if (margin > 32) margin = 32;
  uint32_t KP = k + margin;
  while (KP > k && (uint64_t)KP * G > 16384) --KP;
  return KP;
}

constexpr uint32_t RETRY_CHUNK = 256;  // unverified tensor-pass queries retried per select launch

// bound on |approximate - reference| cosine for the fp32 streaming pass
float eps_stream(uint32_t dim) { return (2.1f * (float)dim + 16.0f) * 5.9604645e-8f; }
// ... and for the streaming pass over the bf16 shadow: the stored rows are rounded to bf16 (2^-9 relative per
// element; Cauchy-Schwarz bounds the sum by the product of the norms), the query stays fp32
float eps_stream_half(uint32_t dim) { return 0.001953125f * 1.02f + eps_stream(dim); }

// Contiguous result block: one D2H copy brings everything the host needs.
struct ResultBlock {
  size_t ok, n, rows, score, dist, ids, total;
  static ResultBlock make(uint64_t B, uint32_t kd) {
    ResultBlock r;
    size_t o = 0;
    r.ok = o;
    o += align_up(B * 4, 16);
    r.n = o;
    o += align_up(B * 4, 16);
    r.rows = o;
    o += align_up(B * kd * 4, 16);
    r.score = o;
    o += align_up(B * kd * 4, 16);
    r.dist = o;
    o += align_up(B * kd * 4, 16);
    r.ids = o;
    o += align_up(B * kd * 16, 16);
    r.total = o;
    return r;
  }
};

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? (T*)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct SearchBufs {
  float* dQ = nullptr;
  float* qnorm = nullptr;
  float* rqnorm = nullptr;
  uint32_t* excl = nullptr;
  uint64_t* cand_keys = nullptr;
  uint16_t* q16 = nullptr;        // normalised bf16 queries (tensor pass)
  uint64_t* lists = nullptr;      // tensor pass private-list scratch
  uint64_t* retry_keys = nullptr; // merged lists for streaming retries of unverified queries
  uint32_t* qmap = nullptr;       // their query indices
  float* dump = nullptr;          // sampled scores for the tensor pass's cut-off bootstrap
  char* res = nullptr;       // ResultBlock (device) when results are workspace-owned
  uint32_t* ok = nullptr;
  uint32_t* n = nullptr;
  uint32_t* rows = nullptr;
  float* score = nullptr;
  float* dist = nullptr;
  uint8_t* ids = nullptr;
  uint32_t* n_total = nullptr;
  uint32_t* totals = nullptr;     // threshold scans: rows that qualify, per query
};

// bound on |approximate - reference| cosine for the bf16 tensor pass: both operands are
// rounded to bf16 (2^-9 relative each, so 2^-8 on every product; Cauchy-Schwarz bounds
// the sum by the product of the norms), plus fp32 accumulation and normalisation slack.
float eps_tensor(uint32_t dim) { return 0.00390625f * 1.02f + ((float)dim + 32.0f) * 1.1920929e-7f; }

// Phases of one tensor-pass scan, as tile counts.  The cut-off a phase works with comes from
// everything scanned before it (first the bootstrap sample of S tiles), so a phase of R tiles
// nominates about KP * R / S_before rows per query; growing the phases geometrically keeps
// that at `growth` * KP per phase (instead of KP * n_tiles / S for a single phase).
std::vector<uint32_t> tensor_phases(uint32_t n_tiles, uint32_t sample_tiles, uint32_t growth, double* hits_per_kp) {
  std::vector<uint32_t> ph;
  double hits = 0.0;
  uint32_t done = 0, seen = sample_tiles ? sample_tiles : 1;
  while (done < n_tiles) {
    uint32_t left = n_tiles - done;
    uint64_t want = growth >= 2 ? (uint64_t)growth * seen : left;
    uint32_t take = (want * 2 >= left) ? left : (uint32_t)want;
    ph.push_back(take);
    hits += (double)take / (double)seen;
    done += take;
    seen = done;
  }
  if (hits_per_kp) *hits_per_kp = hits;
  return ph;
}

// Launch groups of the tensor pass, in query tiles of 128.  A group of g tiles runs as
// g x floor(SMs / g) CTAs, each scanning 1 / floor(SMs / g) of the rows, so some group sizes leave
// SMs idle (128 tiles on 148 SMs: 20 idle for the whole scan).  Splitting the batch into groups of
// SMs, SMs/2 or SMs/4 tiles keeps every SM busy; a small dynamic programme picks the cheapest split
// (cost in units of one CTA scanning the whole shard, plus a little per extra group for its
// bootstrap and launch overhead).
std::vector<uint32_t> tensor_groups(uint64_t n_queries, int sm_count) {
  const uint32_t T = (uint32_t)((n_queries + 127) / 128);
  const uint32_t S = (uint32_t)sm_count;
  std::vector<uint32_t> groups;
  uint32_t left = T;
  while (left >= S) {
    groups.push_back(S);
    left -= S;
  }
  if (left) {
    const uint32_t cand[3] = {S, S / 2, S / 4};
    std::vector<double> cost(left + 1, 0.0);
    std::vector<uint32_t> pick(left + 1, 0);
    for (uint32_t r = 1; r <= left; ++r) {
      cost[r] = 1e30;
      for (int c = -1; c < 3; ++c) {
        const uint32_t g = c < 0 ? r : (cand[c] < r ? cand[c] : r);
        if (!g) continue;
        const double t = 1.0 / (double)(S / g) + 0.02 + cost[r - g];
        if (t < cost[r] - 1e-12) {
          cost[r] = t;
          pick[r] = g;
        }
      }
    }
    for (uint32_t r = left; r;) {
      groups.push_back(pick[r]);
      r -= pick[r];
    }
  }
  // tiles -> queries (the last group takes what is left)
  uint64_t q = n_queries;
  for (auto& g : groups) {
    const uint64_t nq = (uint64_t)g * 128 < q ? (uint64_t)g * 128 : q;
    g = (uint32_t)nq;
    q -= nq;
  }
  return groups;
}

struct Plan {
  uint64_t B;
  uint32_t qlen, ldq, kd;
  uint32_t G, KP;     // streaming pass: producer groups, keys kept per group
  uint32_t KP_wide;   // ... and the wide keep count of the last retry tier (0 = no such tier)
  uint32_t KPt;       // tensor pass: keys kept per query (32 / 64 / 128)
  uint32_t n_slots;   // tensor pass: sampled row tiles for the cut-off bootstrap
  uint32_t growth;    // tensor pass: phase growth factor (tensor_phases)
  uint32_t cap;       // merged-list capacity per query for the primary pass
  uint32_t cap_retry; // ... and for the streaming retries of unverified queries (widest tier)
  uint32_t q_per_launch;  // tensor pass: queries in the largest launch group
  std::vector<uint32_t> groups;  // tensor pass: queries per launch group (tensor_groups)
  bool fast;          // a nominate + rescore pass is usable for this call
  bool tensor;        // ... and it is the tcgen05 pass (else the streaming pass)
  bool half;          // streaming pass over the bf16 shadow (B <= 4)
  bool thr_fast;      // threshold scan served by nominate-all + rescore-all (else the exact path)
  uint32_t thr_cap;   // ... nominee capacity per query
};

// Threshold scans take the fast path when the threshold is far enough above the noise floor
// of random directions that the nominee lists stay short (lower thresholds qualify a large
// part of the corpus: sort-everything on the exact path is the right tool there).
constexpr float THR_FAST_MIN = 0.2f;

Plan make_plan(const cx_index* h, uint64_t B, uint32_t qlen, uint32_t ldq, uint32_t kd, bool threshold_mode,
               float threshold = 0.0f) {
  Plan p;
  p.B = B;
  p.qlen = qlen;
  p.ldq = ldq;
  p.kd = kd;
  const uint32_t n_rows = (uint32_t)h->n_rows;
  p.G = stream_scan_groups(n_rows, h->sm_count);
  p.KP = keep_count(kd, p.G);
  p.cap = p.G * p.KP;
  p.cap_retry = 0;
  p.q_per_launch = 0;
  // rows whose norm under- / overflowed fp32 are outside the fast passes' error bounds (and are never
  // nominated): while the index holds any, everything is served by the exact path
  const bool regular = h->n_irregular() == 0;
  p.fast = regular && !threshold_mode && qlen == h->dim && n_rows >= 256 && kd <= 128 && p.KP >= kd &&
           stream_scan_smem(h->ld, 8, p.KP) != 0 && select_smem(p.cap, h->ld) <= 227 * 1024 &&
           h->force_path != PATH_EXACT;
  p.tensor = false;
  p.half = false;
  p.KPt = 0;
  p.n_slots = 0;
  p.KP_wide = 0;
  p.growth = 0;
  p.thr_fast = false;
  p.thr_cap = 0;
  if (threshold_mode) {
    p.thr_fast = regular && threshold == threshold && threshold >= THR_FAST_MIN && qlen == h->dim && n_rows >= 256 &&
                 h->force_path != PATH_EXACT && stream_scan_smem(h->ld, 8, 64) != 0 &&
                 threshold_rescore_smem(h->ld) <= 200 * 1024;
    if (p.thr_fast) {
      p.KP = 64;  // sizes the streaming pass's shared-memory lists (2 * KP + 128 entries)
      p.thr_cap = B <= 64 ? 8192u : 2048u;
      p.cap = p.thr_cap;
      p.groups = tensor_groups(B, h->sm_count);
      p.q_per_launch = *std::max_element(p.groups.begin(), p.groups.end());
      p.tensor = h->dE16 && h->force_path != PATH_STREAM &&
                 (B >= h->tensor_min_batch || h->force_path == PATH_TENSOR) && tensor_scan_eligible(h->ld16, 16);
    }
    return p;
  }
  if (p.fast) {
    uint32_t w = 4 * p.KP;
    if (w > 256) w = 256;  // the select kernel rescans at most 256 rows
    while (w > p.KP && ((uint64_t)w * p.G > 16384 || stream_scan_smem(h->ld, 4, w) == 0)) w -= 2;
    if (w > p.KP) p.KP_wide = w;
    p.cap_retry = p.G * (p.KP_wide ? p.KP_wide : p.KP);
  }
  if (p.fast && h->dE16 && h->force_path != PATH_STREAM &&
      (B >= h->tensor_min_batch || h->force_path == PATH_TENSOR) && tensor_scan_eligible(h->ld16, kd)) {
    p.tensor = true;
    p.KPt = tensor_keep(kd);
    p.groups = tensor_groups(B, h->sm_count);
    p.q_per_launch = *std::max_element(p.groups.begin(), p.groups.end());
    // the cut-off comes from a sample of S rows, so about KPt * rows / S keys per query clear it;
    // the list capacity is a multiple of that (below; excess is detected by the select kernel and sent to a
    // fallback, never lost silently)
    p.n_slots = tensor_sample_tiles(n_rows, (uint32_t)(B < p.q_per_launch ? B : p.q_per_launch));
    // measured (scripts/k2_probe.py, B = 1024, 1M rows): with a short keep list half the sample does as well
    // and the bootstrap costs half
    if (B > 512 && p.KPt <= 32 && p.n_slots > 16) p.n_slots = 16;
    if (h->tensor_sample_tiles) p.n_slots = std::min<uint32_t>(h->tensor_sample_tiles, tensor_tiles(n_rows));
    double hits_per_kp = 0.0;
    // auto: up to two query tiles with a short keep list the pass is HBM-bound and the epilogue has
    // slack, so one phase (no refine launches) is fastest; otherwise grow x8 (k <= 16) or x6
    p.growth = h->tensor_phase_growth != 0xFFFFFFFFu ? h->tensor_phase_growth
               : (B <= 256 && p.KPt <= 32)           ? 0u
               : (p.KPt <= 32 ? 8u : 6u);
    tensor_phases(tensor_tiles(n_rows), p.n_slots, p.growth, &hits_per_kp);
    const uint64_t expected = (uint64_t)((double)p.KPt * hits_per_kp) + 1;
    // four times the expectation: when a query's KP-th neighbour sits in the dense background of unrelated
    // rows, the 2 * eps the refined cut-off is lowered by admits two to three times the rows the plain
    // order-statistics estimate predicts (measured on 1024-d shards: lists of 2x overflowed for 3 % of queries)
    uint64_t cap = 4 * expected + 4 * p.KPt + 64;
    if (cap > 16384) cap = 16384;
    p.cap = (uint32_t)cap;
  }
  // up to four queries: the streaming pass reads the bf16 shadow (half the bytes of the fp32 rows)
  p.half = p.fast && !p.tensor && B <= 4 && h->dE16 && h->stream_bf16 && h->force_path == PATH_AUTO &&
           stream_scan_half_smem(h->ld16, (uint32_t)B, p.KP) != 0;
  return p;
}

size_t carve_bufs(void* base, const cx_index* h, const Plan& pl, uint32_t n_excl, bool own_queries,
                  bool own_results, SearchBufs* sb) {
  Carver c(base);
  sb->dQ = own_queries ? c.take<float>(pl.B * pl.ldq) : nullptr;
  sb->qnorm = c.take<float>(pl.B);
  sb->rqnorm = c.take<float>(pl.B);
  sb->excl = c.take<uint32_t>(n_excl + 1);
  sb->cand_keys = (pl.fast || pl.thr_fast) ? c.take<uint64_t>((size_t)pl.B * pl.cap) : nullptr;
  sb->totals = c.take<uint32_t>(pl.B);
  if (pl.thr_fast && pl.tensor) {
    sb->q16 = c.take<uint16_t>(align_up(pl.B, 128) * h->ld16);
    sb->lists = (uint64_t*)c.take<char>(tensor_scratch_bytes(h->sm_count));
  }
  if (pl.fast) {
    const uint64_t chunk = pl.B < RETRY_CHUNK ? pl.B : RETRY_CHUNK;
    sb->retry_keys = c.take<uint64_t>((size_t)chunk * pl.cap_retry);
    sb->qmap = c.take<uint32_t>(RETRY_CHUNK);
  }
  if (pl.tensor && pl.fast) {
    sb->q16 = c.take<uint16_t>(align_up(pl.B, 128) * h->ld16);
    sb->lists = (uint64_t*)c.take<char>(tensor_scratch_bytes(h->sm_count));
    const uint64_t nq_launch = pl.B < pl.q_per_launch ? pl.B : pl.q_per_launch;
    sb->dump = c.take<float>(nq_launch * pl.n_slots * 256);
  }
  sb->n_total = c.take<uint32_t>(4);
  const ResultBlock rb = ResultBlock::make(pl.B, pl.kd);
  if (own_results) {
    sb->res = c.take<char>(rb.total);
    if (sb->res) {
      sb->ok = (uint32_t*)(sb->res + rb.ok);
      sb->n = (uint32_t*)(sb->res + rb.n);
      sb->rows = (uint32_t*)(sb->res + rb.rows);
      sb->score = (float*)(sb->res + rb.score);
      sb->dist = (float*)(sb->res + rb.dist);
      sb->ids = (uint8_t*)(sb->res + rb.ids);
    }
  } else {
    sb->ok = c.take<uint32_t>(pl.B);
  }
  return align_up(c.off, 256);
}

// Extras of a pair scan (the dedup self-join): queries are rows of the index itself.
struct PairScan {
  const uint32_t* self_rows = nullptr;  // device [B]: the row each query is
  uint32_t self_base = 0;               // ... which is self_base + b for query b (host-side copy of the same fact)
  bool upper_only = false;              // keep only partners above the query's own row
  uint32_t tile0 = 0;                   // first row tile worth scanning (rows below it cannot qualify)
  // multi-device form: the queries are rows of some shard, the corpus is this shard; a partner qualifies
  // when its global insertion number is above the query's (which also skips the query itself)
  const uint64_t* q_seq = nullptr;      // device [B]
  const uint64_t* h_q_seq = nullptr;    // host copy
};

// phase: RUN_ALL = the whole call; RUN_ENQUEUE = everything up to (and including) the copy of the
// verification flags to the host, no wait (returns CX_PENDING; only for plans with pl.fast);
// RUN_FINISH = wait, then retry / fall back what could not be verified.
enum { RUN_ALL = 0, RUN_ENQUEUE = 1, RUN_FINISH = 2 };
constexpr cx_status CX_PENDING = (cx_status)-1;

// Identity of a device-resident search for graph replay: two calls with equal keys enqueue exactly the
// same kernels with exactly the same arguments.
struct GraphKey {
  uint64_t w[8];
  bool operator==(const uint64_t (&o)[8]) const { return memcmp(w, o, sizeof w) == 0; }
};

uint64_t fnv(uint64_t hsh, const void* p, size_t n) {
  const unsigned char* c = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) hsh = (hsh ^ c[i]) * 0x100000001B3ull;
  return hsh;
}

// every option that shapes the kernels of a search or their arguments (part of a GraphKey)
uint64_t hash_options(uint64_t hsh, const cx_index* h) {
  const uint64_t o[10] = {(uint64_t)h->force_path,           (uint64_t)h->tensor_min_batch,
                          (uint64_t)h->stream_bf16,          (uint64_t)h->tensor_phase_growth,
                          (uint64_t)h->profile,              (uint64_t)h->tensor_sample_tiles,
                          (uint64_t)h->tensor_tune.pair,     (uint64_t)h->tensor_tune.epi_warps,
                          (uint64_t)h->tensor_tune.debug,    (uint64_t)h->tensor_tune.use_leftover_sms};
  return fnv(hsh, o, sizeof o);
}

// Upload of the query batch of a host call, enqueued (and recorded) with the kernels it feeds.
struct HostCopy {
  void* dst;
  const void* src;
  size_t bytes;
};

struct Launcher {
  cx_index* h;
  Workspace* ws;
  cudaStream_t s;
  bool capturing = false;
  uint32_t n_launch = 0;
  cudaError_t record(cudaEvent_t ev) const {
    // inside a capture a plain record would become an internal dependency; the external flag makes
    // it a real timing node that can be read after the graph ran
    return capturing ? cudaEventRecordWithFlags(ev, s, cudaEventRecordExternal) : cudaEventRecord(ev, s);
  }
};

// The fast top-k path, enqueue part: nominate (K1 or K2 in phases) -> select + exact rescore + verify
// -> copy of the verification flags (or of the whole result block) to the host.
cx_status enqueue_topk(Launcher& L, const StoreView& st, const QueryView& qv, const DevFilter& flt, CandView cv,
                       const ResultView& rv, const SearchBufs& sb, const Plan& pl, char* h_block, uint32_t* h_ok,
                       const HostCopy* pre) {
  cx_index* h = L.h;
  cudaStream_t s = L.s;
  const uint64_t B = pl.B;
  const ResultBlock rb = ResultBlock::make(B, pl.kd);
  if (pre && pre->bytes) CU(cudaMemcpyAsync(pre->dst, pre->src, pre->bytes, cudaMemcpyHostToDevice, s));
  if (h->profile) CU(L.record(L.ws->ev0));
  if (pl.tensor) {
    launch_query_bf16(sb.dQ, pl.ldq, h->dim, (uint32_t)B, (uint32_t)align_up(B, 128), sb.q16, h->ld16, s);
    L.n_launch += 1;
    const bool check_rows = flt.has_kinds || flt.has_agent || flt.n_excl || h->n_live != h->n_rows;
    // cut-off bootstrap for every launch group first (the sample buffer is reused in stream order)
    uint64_t q0 = 0;
    for (const uint32_t nq : pl.groups) {
      CU(launch_tensor_bootstrap(st, sb.q16, (uint32_t)q0, nq, flt, check_rows, cv, sb.dump, pl.n_slots, h->sm_count,
                                 s, h->tensor_tune));
      L.n_launch += 2;
      q0 += nq;
    }
    const std::vector<uint32_t> phases = tensor_phases(tensor_tiles(st.n_rows), pl.n_slots, pl.growth, nullptr);
    q0 = 0;
    for (const uint32_t nq : pl.groups) {
      uint32_t tile0 = 0;
      for (size_t ph = 0; ph < phases.size(); ++ph) {
        if (ph) {  // tighten the cut-off with what the earlier phases found
          CU(launch_tau_refine(cv, (uint32_t)q0, nq, 2.0f * eps_tensor(h->dim), s));
          L.n_launch += 1;
        }
        CU(launch_tensor_scan(st, sb.q16, (uint32_t)q0, nq, flt, check_rows, cv, sb.lists, tile0, phases[ph],
                              h->sm_count, s, h->tensor_tune));
        tile0 += phases[ph];
        L.n_launch += 1;
      }
      q0 += nq;
    }
  } else {
    for (uint64_t q0 = 0; q0 < B; q0 += 8) {
      const uint32_t nq = (uint32_t)(B - q0 < 8 ? B - q0 : 8);
      CU(launch_stream_scan(st, qv, (uint32_t)q0, nq, flt, cv, h->sm_count, s, nullptr, nullptr, pl.half));
      L.n_launch += 1;
    }
  }
  if (h->profile) CU(L.record(L.ws->ev1));
  CU(launch_select_rescore(st, qv, 0, (uint32_t)B, cv, rv,
                           pl.tensor ? eps_tensor(h->dim) : pl.half ? eps_stream_half(h->dim) : eps_stream(h->dim),
                           /*scale_by_rqn=*/pl.tensor ? 0 : 1, s));
  L.n_launch += 1;
  if (h_block) CU(cudaMemcpyAsync(h_block, sb.res, rb.total, cudaMemcpyDeviceToHost, s));
  else CU(cudaMemcpyAsync(h_ok, sb.ok, B * 4, cudaMemcpyDeviceToHost, s));
  return CX_OK;
}

void add_pass_time(cx_index* h, Workspace* ws, uint32_t n_pass) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ws->ev0, ws->ev1) == cudaSuccess) {
    h->pass_ns += (uint64_t)(ms * 1e6);
    h->pass_launches += n_pass;
  } else {
    (void)cudaGetLastError();
  }
}

// Core.  Queries are on the device at sb.dQ [B][ldq]; results land in sb.rows/score/
// dist/ids/n (device).  If h_block != nullptr the whole ResultBlock is also copied to
// it (pinned host).  h_ok: pinned host scratch of B words.  Returns with the stream idle.
// gk (optional): identity of the call for graph replay of the enqueue part.
cx_status run_search(cx_index* h, Workspace* ws, const FilterHost& fh, const SearchBufs& sb, const Plan& pl,
                     bool threshold_mode, float threshold, char* h_block, uint32_t* h_ok,
                     uint64_t* h_total /* threshold mode: per-query totals */, const PairScan* tp = nullptr,
                     int phase = RUN_ALL, const GraphKey* gk = nullptr, const HostCopy* pre = nullptr) {
  cudaStream_t s = ws->stream;
  const uint64_t B = pl.B;
  // the query upload of a host call: part of the recorded sequence on the fast top-k path, issued right away otherwise
  const bool pre_in_sequence = pre && pl.fast && !pl.thr_fast && !threshold_mode;
  if (pre && !pre_in_sequence && pre->bytes && phase != RUN_FINISH)
    CU(cudaMemcpyAsync(pre->dst, pre->src, pre->bytes, cudaMemcpyHostToDevice, s));
  const HostCopy* pre_seq = pre_in_sequence ? pre : nullptr;
  StoreView st = h->view();
  QueryView qv;
  qv.Q = sb.dQ;
  qv.qnorm = sb.qnorm;
  qv.rqnorm = sb.rqnorm;
  qv.nq = (uint32_t)B;
  qv.qlen = pl.qlen;
  qv.ldq = pl.ldq;
  DevFilter flt = fh.dev;
  flt.excl_rows = sb.excl;
  flt.n_excl = (uint32_t)fh.excl_rows.size();
  if (flt.n_excl && phase != RUN_FINISH)
    CU(cudaMemcpyAsync(sb.excl, fh.excl_rows.data(), flt.n_excl * 4, cudaMemcpyHostToDevice, s));

  ResultView rv;
  rv.rows = sb.rows;
  rv.score = sb.score;
  rv.dist = sb.dist;
  rv.ids = sb.ids;
  rv.n = sb.n;
  rv.ok = sb.ok;
  rv.k = pl.kd;
  const ResultBlock rb = ResultBlock::make(B, pl.kd);

  if (h->force_path == PATH_STREAM && !pl.fast && !threshold_mode)
    return fail(CX_ERR_VALIDATION, "force_path=stream but the call shape is not eligible");

  std::vector<uint32_t> redo;
  if (pl.thr_fast) {
    // threshold scan: nominate every row within the pass's error bound of the threshold,
    // rescore all nominees exactly, keep `score >= threshold` (index.rs:385)
    CU(ws->ensure_state(B));
    ws->state_dirty = true;
    CandView cv;
    cv.keys = sb.cand_keys;
    cv.cnt = ws->d_cnt;
    cv.gtau = ws->d_gtau;
    cv.cap = pl.cap;
    cv.G = pl.G;
    cv.KP = pl.KP;
    launch_prepare_queries(sb.dQ, sb.qnorm, sb.rqnorm, (uint32_t)B, pl.qlen, pl.ldq, s);
    h->launches += 1;
    // score >= thr implies reference cosine >= thr - 2^-23 (the 1-(1-x) round trip, index.rs:177,255)
    const float slack = 1e-6f;
    uint32_t n_pass = 0;
    if (h->profile) CU(cudaEventRecord(ws->ev0, s));
    if (pl.tensor) {
      const bool check_rows = flt.has_kinds || flt.has_agent || flt.n_excl || h->n_live != h->n_rows;
      launch_query_bf16(sb.dQ, pl.ldq, h->dim, (uint32_t)B, (uint32_t)align_up(B, 128), sb.q16, h->ld16, s);
      cv.KP = 32;
      launch_fill_tau(cv, 0, (uint32_t)B, threshold - eps_tensor(h->dim) - slack, s);
      h->launches += 2;
      const uint32_t t0 = tp ? tp->tile0 : 0, nt = tensor_tiles(st.n_rows) - t0;
      uint64_t q0 = 0;
      for (const uint32_t nq : pl.groups) {
        CU(launch_tensor_scan(st, sb.q16, (uint32_t)q0, nq, flt, check_rows, cv, sb.lists, t0, nt, h->sm_count, s,
                              h->tensor_tune, /*static_tau=*/true));
        ++n_pass;
        q0 += nq;
      }
    } else {
      const float thr_cos = threshold - eps_stream(h->dim) - slack;
      for (uint64_t q0 = 0; q0 < B; q0 += 8) {
        const uint32_t nq = (uint32_t)(B - q0 < 8 ? B - q0 : 8);
        CU(launch_stream_scan(st, qv, (uint32_t)q0, nq, flt, cv, h->sm_count, s, nullptr, &thr_cos));
        ++n_pass;
      }
    }
    if (h->profile) CU(cudaEventRecord(ws->ev1, s));
    CU(launch_threshold_rescore(st, qv, 0, (uint32_t)B, cv, rv, sb.totals, threshold, tp ? tp->self_rows : nullptr,
                                tp ? tp->upper_only : false, s, tp ? tp->q_seq : nullptr, h->dSeq));
    h->launches += n_pass + 1;
    std::vector<uint32_t> tot32;
    if (h_total) tot32.resize(B);
    if (h_block) {
      CU(cudaMemcpyAsync(h_block, sb.res, rb.total, cudaMemcpyDeviceToHost, s));
      h_ok = (uint32_t*)(h_block + rb.ok);
    } else {
      CU(cudaMemcpyAsync(h_ok, sb.ok, B * 4, cudaMemcpyDeviceToHost, s));
    }
    if (h_total) CU(cudaMemcpyAsync(tot32.data(), sb.totals, B * 4, cudaMemcpyDeviceToHost, s));
    CU(ws->wait(h->blocking_sync));
    if (h->profile) add_pass_time(h, ws, n_pass);
    ws->state_dirty = false;
    for (uint64_t b = 0; b < B; ++b) {
      if (!h_ok[b]) redo.push_back((uint32_t)b);
      else if (h_total) h_total[b] = tot32[b];
    }
    (pl.tensor ? h->q_tensor : h->q_stream) += B - redo.size();
    h->fallbacks += redo.size();
    if (redo.empty()) return CX_OK;
    // what overflowed its nominee list (a node with thousands of near-threshold partners) goes to the
    // exact path below, pair rules included
  } else if (pl.fast) {
    CandView cv;
    cv.keys = sb.cand_keys;
    cv.cnt = ws->d_cnt;
    cv.gtau = ws->d_gtau;
    cv.cap = pl.cap;
    cv.G = pl.G;
    cv.KP = pl.tensor ? pl.KPt : pl.KP;
    // scans of the whole shard this call makes (what the profile events cover)
    const uint32_t n_pass = pl.tensor ? (uint32_t)pl.groups.size() : (uint32_t)((B + 7) / 8);
    if (h_block) h_ok = (uint32_t*)(h_block + rb.ok);
    if (phase != RUN_FINISH) {
      CU(ws->ensure_state(B));
      ws->state_dirty = true;  // until the select kernel has re-zeroed it and the stream drained cleanly
      cv.cnt = ws->d_cnt;
      cv.gtau = ws->d_gtau;
      Launcher L{h, ws, s};
      const bool graphable = gk && h->use_graphs && !ws->graph_broken;
      if (graphable && ws->graph && *gk == ws->graph_key) {
        // the same call as the one recorded: one launch replays the whole sequence
        CU(cudaGraphLaunch(ws->graph, s));
        h->graph_launches += 1;
        h->launches += ws->graph_nlaunch;
      } else if (graphable && *gk == ws->last_key) {
        // second time in a row: record the sequence while enqueueing nothing, then launch the recording
        ws->drop_graph();
        cudaGraph_t g = nullptr;
        cx_status est = CX_OK;
        cudaError_t ce = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
        if (ce == cudaSuccess) {
          L.capturing = true;
          est = enqueue_topk(L, st, qv, flt, cv, rv, sb, pl, h_block, h_ok, pre_seq);
          ce = cudaStreamEndCapture(s, &g);
          L.capturing = false;
        }
        if (ce == cudaSuccess && est == CX_OK && g) ce = cudaGraphInstantiate(&ws->graph, g, 0);
        if (g) cudaGraphDestroy(g);
        if (ce != cudaSuccess || est != CX_OK || !ws->graph) {
          // recording is an optimisation: fall back to plain launches for this workspace
          (void)cudaGetLastError();
          ws->drop_graph();
          ws->graph_broken = true;
          Launcher L2{h, ws, s};
          cx_status st2 = enqueue_topk(L2, st, qv, flt, cv, rv, sb, pl, h_block, h_ok, pre_seq);
          if (st2 != CX_OK) return st2;
          h->launches += L2.n_launch;
        } else {
          memcpy(ws->graph_key, gk->w, sizeof ws->graph_key);
          ws->graph_nlaunch = L.n_launch;
          CU(cudaGraphLaunch(ws->graph, s));
          h->graph_launches += 1;
          h->launches += L.n_launch;
        }
      } else {
        if (gk) memcpy(ws->last_key, gk->w, sizeof ws->last_key);
        cx_status est = enqueue_topk(L, st, qv, flt, cv, rv, sb, pl, h_block, h_ok, pre_seq);
        if (est != CX_OK) return est;
        h->launches += L.n_launch;
      }
      if (phase == RUN_ENQUEUE) return CX_PENDING;
    }
    CU(ws->wait(h->blocking_sync));
    if (h->profile) add_pass_time(h, ws, n_pass);
    for (uint64_t b = 0; b < B; ++b)
      if (!h_ok[b]) redo.push_back((uint32_t)b);
    (pl.tensor ? h->q_tensor : h->q_stream) += B - redo.size();
    if (pl.half) h->q_stream16 += B - redo.size();
    h->fallbacks += redo.size();
    // Retry ladder for queries the primary pass could not verify: the fp32 streaming pass
    // (error bound ~100x tighter than the bf16 pass), four queries per pass over the matrix,
    // first with the normal keep count, then with a wide one (near-ties around rank k need
    // more rescored rows, not a different algorithm).  What still fails goes to the exact path.
    const uint32_t ladder[2] = {(pl.tensor || pl.half) ? pl.KP : 0u, pl.KP_wide > pl.KP ? pl.KP_wide : 0u};
    bool retried = false;
    for (int tier = 0; tier < 2 && !redo.empty(); ++tier) {
      const uint32_t KPr = ladder[tier];
      if (!KPr) continue;
      retried = true;
      CandView cr = cv;
      cr.cap = pl.G * KPr;
      cr.G = pl.G;
      cr.KP = KPr;
      cr.q_base = 0;
      for (size_t c0 = 0; c0 < redo.size(); c0 += RETRY_CHUNK) {
        const uint32_t nc = (uint32_t)std::min<size_t>(RETRY_CHUNK, redo.size() - c0);
        CU(cudaMemcpyAsync(sb.qmap, redo.data() + c0, nc * 4, cudaMemcpyHostToDevice, s));
        // queries per pass over the matrix: eight with the normal keep count, four with the wide one (its lists
        // are four times as long and share the same shared memory)
        const uint32_t per_pass = (tier == 0 && stream_scan_smem(h->ld, 8, KPr) != 0) ? 8u : 4u;
        for (uint32_t g0 = 0; g0 < nc; g0 += per_pass) {
          const uint32_t ng = nc - g0 < per_pass ? nc - g0 : per_pass;
          cr.keys = sb.retry_keys + (size_t)g0 * cr.cap;
          CU(launch_stream_scan(st, qv, 0, ng, flt, cr, h->sm_count, s, sb.qmap + g0));
          h->launches += 1;
        }
        cr.keys = sb.retry_keys;
        CU(launch_select_rescore(st, qv, 0, nc, cr, rv, eps_stream(h->dim), 1, s, sb.qmap));
        h->launches += 1;
      }
      CU(cudaMemcpyAsync(h_ok, sb.ok, B * 4, cudaMemcpyDeviceToHost, s));
      CU(ws->wait(h->blocking_sync));
      std::vector<uint32_t> still;
      for (uint32_t b : redo)
        if (!h_ok[b]) still.push_back(b);
      h->q_stream += redo.size() - still.size();
      redo.swap(still);
    }
    if (retried && redo.empty()) {
      if (h_block) CU(cudaMemcpyAsync(h_block, sb.res, rb.total, cudaMemcpyDeviceToHost, s));
      CU(ws->wait(h->blocking_sync));
    }
    ws->state_dirty = false;
    if (redo.empty()) return CX_OK;
  } else {
    launch_prepare_queries(sb.dQ, sb.qnorm, sb.rqnorm, (uint32_t)B, pl.qlen, pl.ldq, s);
    h->launches += 1;
    redo.resize(B);
    for (uint64_t b = 0; b < B; ++b) redo[b] = (uint32_t)b;
  }

  // exact path: reference arithmetic for every row, radix sort, emit.  Its buffers (two key arrays over
  // the whole shard + sort scratch) belong to the workspace and exist only once a query got here.
  uint32_t min_ord = 1;  // every real key, NaN included (they sort last)
  if (threshold_mode) {
    if (threshold != threshold) min_ord = 0xFFFFFFFFu;  // score >= NaN is false
    else min_ord = ord_from_score(threshold > 0.0f ? threshold : 0.0f);
  }
  const size_t keys_bytes = align_up((size_t)st.n_rows * 8, 256), sort_bytes = exact_sort_tmp_bytes(st.n_rows);
  CU(ws->ensure_exact(2 * keys_bytes + sort_bytes));
  uint64_t* ekeys_a = (uint64_t*)ws->dx;
  uint64_t* ekeys_b = (uint64_t*)((char*)ws->dx + keys_bytes);
  void* sort_tmp = (char*)ws->dx + 2 * keys_bytes;
  for (uint32_t b : redo) {
    if (tp && tp->q_seq)
      launch_exact_keys(st, qv, b, flt, ekeys_a, s, 0xFFFFFFFFu, false, h->dSeq, tp->h_q_seq[b]);
    else
      launch_exact_keys(st, qv, b, flt, ekeys_a, s, tp ? tp->self_base + b : 0xFFFFFFFFu, tp ? tp->upper_only : false);
    CU(exact_sort(ekeys_a, ekeys_b, st.n_rows, sort_tmp, sort_bytes, s));
    launch_exact_emit(st, qv, b, ekeys_b, st.n_rows, pl.kd, min_ord, sb.rows + (size_t)b * pl.kd,
                      sb.score + (size_t)b * pl.kd, sb.dist + (size_t)b * pl.kd,
                      sb.ids ? sb.ids + (size_t)b * pl.kd * 16 : nullptr, sb.n + b, sb.n_total, s);
    h->launches += 4;
    if (h_total) {
      uint32_t t32 = 0;
      CU(cudaMemcpyAsync(&t32, sb.n_total, 4, cudaMemcpyDeviceToHost, s));
      CU(ws->wait(h->blocking_sync));
      h_total[b] = t32;
    }
  }
  h->q_exact += redo.size();
  CU(cudaGetLastError());
  if (h_block) CU(cudaMemcpyAsync(h_block, sb.res, rb.total, cudaMemcpyDeviceToHost, s));
  CU(ws->wait(h->blocking_sync));
  return CX_OK;
}

// Device part of a threshold scan whose queries already sit in device memory (used by the host
// entry points, the dedup self-join and the multi-device index).  Results [B][kd] land in the
// workspace-owned block (sb.*), per-query totals in totals[].  Returns with the stream idle.
struct ThresholdCall {
  WsLease lease;
  FilterHost fh;
  SearchBufs sb;
  Plan pl;
  explicit ThresholdCall(cx_index* h) : lease(h) {}
};

}  // namespace

namespace cx {

// Batched search_threshold on one device with device-resident queries [B][dim] and device-resident
// results: rows / score / dist [B][cap] (+ ids if d_ids), n [B] entries written, d_total [B] rows that
// qualify (all on this index's device).  pair (optional) applies the dedup scanner's pair rules.
// Blocks until the results are complete.
cx_status index_threshold_device(cx_index* h, const float* d_queries, uint32_t qlen, uint32_t q_stride, uint64_t B,
                                 float threshold, const cx_filter* filter, uint64_t cap, uint32_t* d_rows,
                                 float* d_score, float* d_dist, uint8_t* d_ids, uint32_t* d_n, uint32_t* d_total,
                                 uint64_t* h_total, const PairRules* pair) {
  if (!B) return CX_OK;
  CU(cudaSetDevice(h->device));
  cx_status st = settle(h);
  if (st != CX_OK) return st;
  const uint64_t kd64 = cap < h->n_rows ? cap : h->n_rows;
  ThresholdCall t(h);
  CU(t.lease.init());
  Workspace* ws = t.lease.ws;
  if (h_total)
    for (uint64_t b = 0; b < B; ++b) h_total[b] = 0;
  if (h->n_live == 0 || kd64 == 0) {
    CU(cudaMemsetAsync(d_n, 0, B * 4, ws->stream));
    if (d_total) CU(cudaMemsetAsync(d_total, 0, B * 4, ws->stream));
    CU(ws->wait(h->blocking_sync));
    return CX_OK;
  }
  const uint32_t kd = (uint32_t)kd64;
  const uint32_t ldq = (uint32_t)align_up(qlen > h->ld ? qlen : h->ld, 4);
  build_filter(h, filter, &t.fh);
  t.pl = make_plan(h, B, qlen, ldq, kd, true, threshold);
  const bool own_q = q_stride != ldq;  // rows of the store itself are already padded to ld
  const uint32_t n_excl = (uint32_t)t.fh.excl_rows.size();
  const bool direct_out = kd == cap;  // the caller's [B][cap] layout is the kernels' [B][kd] layout
  const size_t dbytes = carve_bufs(nullptr, h, t.pl, n_excl, own_q, !direct_out, &t.sb);
  CU(ws->ensure(dbytes, align_up(B * 4, 256)));
  carve_bufs(ws->d, h, t.pl, n_excl, own_q, !direct_out, &t.sb);
  SearchBufs& sb = t.sb;
  cudaStream_t s = ws->stream;
  if (own_q) {
    CU(cudaMemsetAsync(sb.dQ, 0, B * ldq * 4, s));
    CU(cudaMemcpy2DAsync(sb.dQ, ldq * 4, d_queries, (size_t)q_stride * 4, (size_t)qlen * 4, B, cudaMemcpyDeviceToDevice, s));
  } else {
    sb.dQ = const_cast<float*>(d_queries);
  }
  if (direct_out) {
    sb.rows = d_rows;
    sb.score = d_score;
    sb.dist = d_dist;
    sb.ids = d_ids;
    sb.n = d_n;
  }
  std::vector<uint64_t> totals(B, 0);
  PairScan tp;
  if (pair) {
    tp.self_rows = pair->d_self_rows;
    tp.self_base = pair->self_base;
    tp.upper_only = pair->upper_only;
    tp.tile0 = pair->first_row / 256;
    tp.q_seq = pair->d_q_seq;
    tp.h_q_seq = pair->h_q_seq;
  }
  st = run_search(h, ws, t.fh, sb, t.pl, true, threshold, nullptr, (uint32_t*)ws->hp, totals.data(),
                  pair ? &tp : nullptr);
  if (st != CX_OK) return st;
  if (!direct_out) {  // cap > rows in the shard: spread [B][kd] into the caller's [B][cap]
    CU(cudaMemcpy2DAsync(d_rows, cap * 4, sb.rows, kd * 4, kd * 4, B, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpy2DAsync(d_score, cap * 4, sb.score, kd * 4, kd * 4, B, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpy2DAsync(d_dist, cap * 4, sb.dist, kd * 4, kd * 4, B, cudaMemcpyDeviceToDevice, s));
    if (d_ids) CU(cudaMemcpy2DAsync(d_ids, cap * 16, sb.ids, kd * 16, kd * 16, B, cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(d_n, sb.n, B * 4, cudaMemcpyDeviceToDevice, s));
  }
  if (h_total)
    for (uint64_t b = 0; b < B; ++b) h_total[b] = totals[b];
  if (d_total) {
    uint32_t* ht = (uint32_t*)ws->hp;
    for (uint64_t b = 0; b < B; ++b) ht[b] = (uint32_t)totals[b];
    CU(cudaMemcpyAsync(d_total, ht, B * 4, cudaMemcpyHostToDevice, s));
  }
  CU(ws->wait(h->blocking_sync));
  return CX_OK;
}

}  // namespace cx

namespace {

cx_status search_host(cx_index* h, const float* queries, uint64_t B, uint32_t qlen, uint64_t k,
                      const cx_filter* filter, bool threshold_mode, float threshold, uint8_t* out_ids,
                      float* out_score, float* out_dist, uint64_t* out_n, uint64_t* out_total) {
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (B && !queries) return fail(CX_ERR_VALIDATION, "null queries");
  if (!out_n) return fail(CX_ERR_VALIDATION, "null out_n");
  for (uint64_t b = 0; b < B; ++b) out_n[b] = 0;
  if (out_total)
    for (uint64_t b = 0; b < B; ++b) out_total[b] = 0;
  if (h->shards)
    return shard_search_host(h, queries, B, qlen, k, filter, threshold_mode, threshold, out_ids, out_score, out_dist,
                             out_n, out_total);
  if (h->n_live == 0 || B == 0) return CX_OK;  // index.rs:331: empty -> Ok(vec![])
  CU(cudaSetDevice(h->device));
  cx_status sst = settle(h);
  if (sst != CX_OK) return sst;
  const uint64_t kd64 = k < h->n_rows ? k : h->n_rows;
  if (kd64 == 0) return CX_OK;
  const uint32_t kd = (uint32_t)kd64;
  const uint32_t ldq = (uint32_t)align_up(qlen > h->ld ? qlen : h->ld, 4);

  FilterHost fh;
  build_filter(h, filter, &fh);
  const Plan pl = make_plan(h, B, qlen, ldq, kd, threshold_mode, threshold);

  WsLease lease(h);
  CU(lease.init());
  Workspace* ws = lease.ws;
  SearchBufs sb;
  const size_t dbytes = carve_bufs(nullptr, h, pl, (uint32_t)fh.excl_rows.size(), true, true, &sb);
  const ResultBlock rb = ResultBlock::make(B, kd);
  const size_t hq = align_up(B * ldq * 4, 256);
  CU(ws->ensure(dbytes, hq + rb.total));
  carve_bufs(ws->d, h, pl, (uint32_t)fh.excl_rows.size(), true, true, &sb);
  char* hp = (char*)ws->hp;
  float* hQ = (float*)hp;
  char* h_block = hp + hq;

  cudaStream_t s = ws->stream;
  bool direct = false;
  if (qlen == ldq) {
    // rows need no padding: if the caller's buffer is pinned, DMA straight from it (pageable
    // memory goes through the staging copy: the driver's own pageable path serialises badly)
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, queries) == cudaSuccess) direct = pa.type == cudaMemoryTypeHost;
    else (void)cudaGetLastError();
  }
  HostCopy pre{sb.dQ, queries, (size_t)B * ldq * 4};
  if (!direct) {
    // stage queries, zero padded to ldq
    if (qlen == ldq) {
      memcpy(hQ, queries, (size_t)B * qlen * 4);
    } else {
      for (uint64_t b = 0; b < B; ++b) {
        memcpy(hQ + b * ldq, queries + b * qlen, (size_t)qlen * 4);
        for (uint32_t d = qlen; d < ldq; ++d) hQ[b * ldq + d] = 0.0f;
      }
    }
    pre.src = hQ;
  }
  h->h2d += B * ldq * 4;
  (void)s;

  // a repeated call shape (the interactive B = 1 search, a fixed-size search_batch) replays upload + kernels +
  // download as ONE graph launch
  GraphKey gk;
  const bool graphable = !threshold_mode && pl.fast && fh.excl_rows.empty();
  if (graphable) {
    const DevFilter& f = fh.dev;
    uint64_t hs = 0xCBF29CE484222325ull;
    hs = fnv(hs, f.kind_mask, sizeof f.kind_mask);
    hs = fnv(hs, &f.agent, sizeof f.agent);
    hs = fnv(hs, &f.has_kinds, sizeof f.has_kinds);
    hs = fnv(hs, &f.has_agent, sizeof f.has_agent);
    const uint64_t misc[3] = {(uint64_t)(uintptr_t)h_block, (uint64_t)(uintptr_t)h->dE, 0x486f7374ull /* host call */};
    hs = fnv(hs, misc, sizeof misc);
    hs = hash_options(hs, h);
    gk.w[0] = B;
    gk.w[1] = kd | ((uint64_t)qlen << 32);
    gk.w[2] = h->n_rows;
    gk.w[3] = h->n_live;
    gk.w[4] = (uint64_t)(uintptr_t)pre.src;
    gk.w[5] = (uint64_t)(uintptr_t)ws->d;
    gk.w[6] = (uint64_t)(uintptr_t)ws->hp;
    gk.w[7] = hs | 1ull;
  }

  std::vector<uint64_t> totals;
  if (threshold_mode) totals.assign(B, 0);
  cx_status stt = run_search(h, ws, fh, sb, pl, threshold_mode, threshold, h_block, nullptr,
                             threshold_mode ? totals.data() : nullptr, nullptr, RUN_ALL, graphable ? &gk : nullptr, &pre);
  if (stt != CX_OK) return stt;
  h->d2h += rb.total;

  const uint32_t* h_n = (const uint32_t*)(h_block + rb.n);
  const float* h_score = (const float*)(h_block + rb.score);
  const float* h_dist = (const float*)(h_block + rb.dist);
  const uint8_t* h_ids = (const uint8_t*)(h_block + rb.ids);
  if (kd == k) {  // same layout on both sides: three block copies instead of 3B row copies
    for (uint64_t b = 0; b < B; ++b) {
      out_n[b] = h_n[b];
      if (out_total) out_total[b] = threshold_mode ? totals[b] : h_n[b];
    }
    if (out_score) memcpy(out_score, h_score, (size_t)B * kd * 4);
    if (out_dist) memcpy(out_dist, h_dist, (size_t)B * kd * 4);
    if (out_ids) memcpy(out_ids, h_ids, (size_t)B * kd * 16);
    return CX_OK;
  }
  for (uint64_t b = 0; b < B; ++b) {
    const uint32_t n = h_n[b];
    out_n[b] = n;
    if (out_total) out_total[b] = threshold_mode ? totals[b] : n;
    if (out_score) memcpy(out_score + b * k, h_score + b * kd, (size_t)n * 4);
    if (out_dist) memcpy(out_dist + b * k, h_dist + b * kd, (size_t)n * 4);
    if (out_ids) memcpy(out_ids + b * k * 16, h_ids + b * kd * 16, (size_t)n * 16);
  }
  return CX_OK;
}

}  // namespace

// Test hook (no device needed): the launch plan of the tensor pass for a batch of n_queries against
// n_rows rows on a GPU with sm_count SMs.  groups[] = queries per launch group (tensor_groups),
// phases[] = row tiles per scan phase (tensor_phases with `sample_tiles` bootstrap tiles and the
// given growth); *_n in: capacity, out: entries written.
extern "C" cx_status cx_debug_tensor_plan(uint64_t n_queries, uint64_t n_rows, int sm_count, uint32_t sample_tiles,
                                          uint32_t growth, uint32_t* groups, uint32_t* groups_n, uint32_t* phases,
                                          uint32_t* phases_n, double* hits_per_kp) {
  if (!groups || !groups_n || !phases || !phases_n || sm_count < 4) return fail(CX_ERR_VALIDATION, "bad argument");
  const std::vector<uint32_t> g = tensor_groups(n_queries, sm_count);
  double hk = 0.0;
  const std::vector<uint32_t> ph = tensor_phases((uint32_t)((n_rows + 255) / 256), sample_tiles, growth, &hk);
  if (g.size() > *groups_n || ph.size() > *phases_n) return fail(CX_ERR_VALIDATION, "output too small");
  std::copy(g.begin(), g.end(), groups);
  std::copy(ph.begin(), ph.end(), phases);
  *groups_n = (uint32_t)g.size();
  *phases_n = (uint32_t)ph.size();
  if (hits_per_kp) *hits_per_kp = hk;
  return CX_OK;
}

extern "C" cx_status cx_search(cx_index* h, const float* query, uint32_t qlen, uint64_t k,
                               const cx_filter* filter, uint8_t* out_ids, float* out_score, float* out_distance,
                               uint64_t* out_n) {
  cx::CallerDevice keep_callers_device;
  return search_host(h, query, 1, qlen, k, filter, false, 0.0f, out_ids, out_score, out_distance, out_n, nullptr);
}

extern "C" cx_status cx_search_batch(cx_index* h, const float* queries, uint64_t B, uint32_t qlen, uint64_t k,
                                     const cx_filter* filter, uint8_t* out_ids, float* out_score,
                                     float* out_distance, uint64_t* out_n) {
  cx::CallerDevice keep_callers_device;
  return search_host(h, queries, B, qlen, k, filter, false, 0.0f, out_ids, out_score, out_distance, out_n,
                     nullptr);
}

extern "C" cx_status cx_search_threshold(cx_index* h, const float* query, uint32_t qlen, float threshold,
                                         const cx_filter* filter, uint64_t cap, uint8_t* out_ids,
                                         float* out_score, float* out_distance, uint64_t* out_n,
                                         uint64_t* out_total) {
  cx::CallerDevice keep_callers_device;
  uint64_t total = 0;
  cx_status st = search_host(h, query, 1, qlen, cap, filter, true, threshold, out_ids, out_score, out_distance,
                             out_n, &total);
  if (out_total) *out_total = total;
  return st;
}

extern "C" cx_status cx_search_threshold_batch(cx_index* h, const float* queries, uint64_t B, uint32_t qlen,
                                               float threshold, const cx_filter* filter, uint64_t cap,
                                               uint8_t* out_ids, float* out_score, float* out_distance,
                                               uint64_t* out_n, uint64_t* out_total) {
  cx::CallerDevice keep_callers_device;
  return search_host(h, queries, B, qlen, cap, filter, true, threshold, out_ids, out_score, out_distance, out_n,
                     out_total);
}

// ------------------------------------------------------------------------------
// DedupScanner::scan (linker/dedup.rs:65-127) restricted to what the similarity scan decides:
// every live row searches the index with search_threshold(embedding, threshold) (:84-86),
// skips itself (:91-93) and reports each unordered pair once (:96-105).  Here the rows of the
// index are their own query batch (device resident, no copy), the tcgen05 pass scans only the
// row tiles at or above the batch (a pair is reported from its lower row), and the exact
// rescoring keeps `score >= threshold` among partners above the query's own row.
// Pairs come out ordered by (row of a asc, score desc, row of b asc) -- the order the reference
// produces when storage lists nodes in insertion order.  The per-node lists are compacted on the
// device (prefix sum over the per-node counts, dense (a, b, score) triples): the copy back is the
// size of the output, not of the [nodes][per_node_cap] result block.
extern "C" cx_status cx_dedup_scan(cx_index* h, float threshold, uint32_t per_node_cap, uint64_t max_pairs,
                                   uint8_t* out_a_ids, uint8_t* out_b_ids, float* out_score, uint64_t* out_n,
                                   uint64_t* out_total) {
  cx::CallerDevice keep_callers_device;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (!out_n) return fail(CX_ERR_VALIDATION, "null out_n");
  *out_n = 0;
  if (out_total) *out_total = 0;
  if (max_pairs && (!out_a_ids || !out_b_ids || !out_score)) return fail(CX_ERR_VALIDATION, "null output buffer");
  if (h->shards)
    return shard_dedup_scan(h, threshold, per_node_cap, max_pairs, out_a_ids, out_b_ids, out_score, out_n, out_total);
  if (h->n_live < 2 || !per_node_cap || threshold != threshold) return CX_OK;
  CU(cudaSetDevice(h->device));
  cx_status sst = settle(h);
  if (sst != CX_OK) return sst;
  const uint32_t n_rows = (uint32_t)h->n_rows;
  const uint32_t kd = per_node_cap < n_rows ? per_node_cap : n_rows;
  const uint64_t QB = (uint64_t)h->sm_count * 128u;  // one tensor-pass launch group
  uint64_t n_out = 0, n_tot = 0;
  // outer workspace: self rows, per-node results, the compacted pairs and their pinned mirror
  WsLease outer(h);
  CU(outer.init());
  Workspace* wo = outer.ws;
  const uint64_t Bmax = n_rows < QB ? n_rows : QB;
  const size_t o_self = 0, o_rows = align_up(o_self + Bmax * 4, 256), o_sc = align_up(o_rows + Bmax * kd * 4, 256),
               o_di = align_up(o_sc + Bmax * kd * 4, 256), o_n = align_up(o_di + Bmax * kd * 4, 256),
               o_tot = align_up(o_n + Bmax * 4, 256), o_off = align_up(o_tot + Bmax * 4, 256),
               o_pairs = align_up(o_off + (Bmax + 1) * 4, 256);
  const uint64_t pair_cap = std::min<uint64_t>(Bmax * kd, max_pairs ? max_pairs : 1);
  const size_t dev_bytes = o_pairs + pair_cap * 12;
  const size_t host_bytes = align_up(pair_cap * 12, 256) + align_up(Bmax * 4, 256) + 256;
  CU(wo->ensure_aux(dev_bytes, host_bytes));
  char* d = (char*)wo->aux;
  char* hpin = (char*)wo->aux_h;
  uint32_t* h_pairs = (uint32_t*)hpin;
  uint32_t* h_tot = (uint32_t*)(hpin + align_up(pair_cap * 12, 256));
  uint32_t* h_cnt = (uint32_t*)(hpin + align_up(pair_cap * 12, 256) + align_up(Bmax * 4, 256));
  for (uint64_t r0 = 0; r0 < n_rows; r0 += QB) {
    const uint64_t B = n_rows - r0 < QB ? n_rows - r0 : QB;
    launch_iota(reinterpret_cast<uint32_t*>(d + o_self), (uint32_t)B, (uint32_t)r0, wo->stream);
    CU(cudaStreamSynchronize(wo->stream));
    PairRules pr;
    pr.d_self_rows = reinterpret_cast<uint32_t*>(d + o_self);
    pr.self_base = (uint32_t)r0;
    pr.upper_only = true;
    pr.first_row = (uint32_t)r0;
    cx_status stt = index_threshold_device(h, h->dE + (size_t)r0 * h->ld, h->dim, h->ld, B, threshold, nullptr, kd,
                                           reinterpret_cast<uint32_t*>(d + o_rows), reinterpret_cast<float*>(d + o_sc),
                                           reinterpret_cast<float*>(d + o_di), nullptr,
                                           reinterpret_cast<uint32_t*>(d + o_n), reinterpret_cast<uint32_t*>(d + o_tot),
                                           nullptr, &pr);
    if (stt != CX_OK) return stt;
    // removed nodes search for nothing (dedup.rs:73-76): their counts are zeroed by the compaction
    const uint64_t room = max_pairs - n_out;
    launch_compact_pairs(reinterpret_cast<uint32_t*>(d + o_rows), reinterpret_cast<float*>(d + o_sc),
                         reinterpret_cast<uint32_t*>(d + o_n), reinterpret_cast<uint32_t*>(d + o_tot), h->dMeta,
                         (uint32_t)r0, (uint32_t)B, kd, reinterpret_cast<uint32_t*>(d + o_off),
                         reinterpret_cast<uint32_t*>(d + o_pairs), (uint32_t)std::min<uint64_t>(room, pair_cap),
                         wo->stream);
    h->launches += 3;
    CU(cudaMemcpyAsync(h_cnt, d + o_off + B * 4, 4, cudaMemcpyDeviceToHost, wo->stream));
    CU(cudaMemcpyAsync(h_tot, d + o_tot, B * 4, cudaMemcpyDeviceToHost, wo->stream));
    CU(cudaStreamSynchronize(wo->stream));
    const uint64_t got = std::min<uint64_t>(*h_cnt, std::min<uint64_t>(room, pair_cap));
    for (uint64_t i = 0; i < B; ++i) n_tot += h_tot[i];  // already zero for removed nodes
    if (got) {
      CU(cudaMemcpyAsync(h_pairs, d + o_pairs, got * 12, cudaMemcpyDeviceToHost, wo->stream));
      CU(cudaStreamSynchronize(wo->stream));
      h->d2h += got * 12;
      for (uint64_t j = 0; j < got; ++j, ++n_out) {
        memcpy(out_a_ids + 16 * n_out, h->h_ids.data() + 16 * (size_t)h_pairs[3 * j], 16);
        memcpy(out_b_ids + 16 * n_out, h->h_ids.data() + 16 * (size_t)h_pairs[3 * j + 1], 16);
        memcpy(out_score + n_out, &h_pairs[3 * j + 2], 4);
      }
    }
    h->d2h += B * 4 + 4;
  }
  *out_n = n_out;
  if (out_total) *out_total = n_tot;
  return CX_OK;
}

// A device-resident batch search in flight (cx_search_batch_device_begin .. _end).
struct DeviceSearch {
  WsLease lease;
  FilterHost fh;
  SearchBufs sb;
  Plan pl;
  explicit DeviceSearch(cx_index* h) : lease(h) {}
};

// Shared by the one-call and the begin / end forms.  ticket == nullptr: run to completion.
cx_status cx::index_search_device(cx_index* h, const float* d_queries, uint32_t qlen, uint64_t B, uint64_t k,
                                  const cx_filter* filter, uint32_t* d_out_rows, float* d_out_score,
                                  float* d_out_distance, uint8_t* d_out_ids, uint32_t* d_out_n, void* stream,
                                  void** ticket) {
  if (ticket) *ticket = nullptr;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (!d_queries || !d_out_rows || !d_out_score || !d_out_distance || !d_out_n)
    return fail(CX_ERR_VALIDATION, "null device buffer");
  if (B == 0) return CX_OK;
  CU(cudaSetDevice(h->device));
  cx_status sst = settle(h);
  if (sst != CX_OK) return sst;
  cudaStream_t user = (cudaStream_t)stream;
  if (h->n_live == 0 || k == 0) {
    CU(cudaMemsetAsync(d_out_n, 0, B * 4, user));
    return CX_OK;
  }
  if (k > h->n_rows) return fail(CX_ERR_VALIDATION, "device search needs k <= rows in the shard");
  const uint32_t kd = (uint32_t)k;
  const uint32_t ldq = (uint32_t)align_up(qlen > h->ld ? qlen : h->ld, 4);
  std::unique_ptr<DeviceSearch> t(new DeviceSearch(h));
  build_filter(h, filter, &t->fh);
  t->pl = make_plan(h, B, qlen, ldq, kd, false);
  CU(t->lease.init());
  Workspace* ws = t->lease.ws;
  const bool own_q = qlen != ldq;
  const size_t dbytes = carve_bufs(nullptr, h, t->pl, (uint32_t)t->fh.excl_rows.size(), own_q, false, &t->sb);
  CU(ws->ensure(dbytes, align_up(B * 4, 256)));
  carve_bufs(ws->d, h, t->pl, (uint32_t)t->fh.excl_rows.size(), own_q, false, &t->sb);
  SearchBufs& sb = t->sb;
  // order after whatever produced the queries on the caller's stream
  CU(cudaEventRecord(ws->ev_sync, user));
  CU(cudaStreamWaitEvent(ws->stream, ws->ev_sync, 0));
  if (own_q) {
    CU(cudaMemsetAsync(sb.dQ, 0, B * ldq * 4, ws->stream));
    CU(cudaMemcpy2DAsync(sb.dQ, ldq * 4, d_queries, (size_t)qlen * 4, (size_t)qlen * 4, B, cudaMemcpyDeviceToDevice,
                         ws->stream));
  } else {
    sb.dQ = const_cast<float*>(d_queries);
  }
  sb.rows = d_out_rows;
  sb.score = d_out_score;
  sb.dist = d_out_distance;
  sb.ids = d_out_ids;
  sb.n = d_out_n;
  // identity of this call for graph replay: everything that shapes the kernels or their arguments
  GraphKey gk;
  {
    const DevFilter& f = t->fh.dev;
    uint64_t hs = 0xCBF29CE484222325ull;
    hs = fnv(hs, f.kind_mask, sizeof f.kind_mask);
    hs = fnv(hs, &f.agent, sizeof f.agent);
    hs = fnv(hs, &f.has_kinds, sizeof f.has_kinds);
    hs = fnv(hs, &f.has_agent, sizeof f.has_agent);
    const uint64_t misc[5] = {(uint64_t)(uintptr_t)d_out_distance, (uint64_t)(uintptr_t)d_out_ids,
                              (uint64_t)(uintptr_t)d_out_n, (uint64_t)(uintptr_t)ws->hp,
                              (uint64_t)t->fh.excl_rows.size()};
    hs = fnv(hs, misc, sizeof misc);
    hs = hash_options(hs, h);
    gk.w[0] = B;
    gk.w[1] = kd | ((uint64_t)qlen << 32);
    gk.w[2] = h->n_rows;
    gk.w[3] = h->n_live;
    gk.w[4] = (uint64_t)(uintptr_t)d_queries;
    gk.w[5] = (uint64_t)(uintptr_t)d_out_rows;
    gk.w[6] = (uint64_t)(uintptr_t)d_out_score ^ ((uint64_t)(uintptr_t)ws->d << 1) ^ ((uint64_t)(uintptr_t)h->dE >> 3);
    gk.w[7] = hs | 1ull;  // never all-zero (an empty key slot)
  }
  const bool split = ticket != nullptr && t->pl.fast;
  cx_status st = run_search(h, ws, t->fh, sb, t->pl, false, 0.0f, nullptr, (uint32_t*)ws->hp, nullptr, nullptr,
                            split ? RUN_ENQUEUE : RUN_ALL, &gk);
  if (st == CX_PENDING) {
    // the caller's stream continues after the results (they are final unless _end reports retries)
    CU(cudaEventRecord(ws->ev_sync, ws->stream));
    CU(cudaStreamWaitEvent(user, ws->ev_sync, 0));
    *ticket = t.release();
    return CX_OK;
  }
  // run_search returned with its stream idle: the results are complete and visible
  return st;
}

extern "C" cx_status cx_search_batch_device(cx_index* h, const float* d_queries, uint64_t B, uint64_t k,
                                            const cx_filter* filter, uint32_t* d_out_rows, float* d_out_score,
                                            float* d_out_distance, uint8_t* d_out_ids, uint32_t* d_out_n,
                                            void* stream) {
  cx::CallerDevice keep_callers_device;
  if (h && h->shards)
    return shard_search_device(h, d_queries, B, k, filter, d_out_rows, d_out_score, d_out_distance, d_out_ids, d_out_n,
                               stream, nullptr);
  return index_search_device(h, d_queries, h ? h->dim : 0, B, k, filter, d_out_rows, d_out_score, d_out_distance,
                             d_out_ids, d_out_n, stream, nullptr);
}

extern "C" cx_status cx_search_batch_device_begin(cx_index* h, const float* d_queries, uint64_t B, uint64_t k,
                                                  const cx_filter* filter, uint32_t* d_out_rows,
                                                  float* d_out_score, float* d_out_distance, uint8_t* d_out_ids,
                                                  uint32_t* d_out_n, void* stream, void** ticket) {
  cx::CallerDevice keep_callers_device;
  if (!ticket) return fail(CX_ERR_VALIDATION, "null ticket");
  if (h && h->shards)
    return shard_search_device(h, d_queries, B, k, filter, d_out_rows, d_out_score, d_out_distance, d_out_ids, d_out_n,
                               stream, ticket);
  return index_search_device(h, d_queries, h ? h->dim : 0, B, k, filter, d_out_rows, d_out_score, d_out_distance,
                             d_out_ids, d_out_n, stream, ticket);
}

extern "C" cx_status cx_search_ticket_ok(void* ticket, const uint32_t** d_ok) {
  if (!d_ok) return fail(CX_ERR_VALIDATION, "null argument");
  *d_ok = ticket ? ((DeviceSearch*)ticket)->sb.ok : nullptr;
  return CX_OK;
}

extern "C" cx_status cx_search_batch_device_end(cx_index* h, void* ticket, uint64_t* n_redone) {
  cx::CallerDevice keep_callers_device;
  if (n_redone) *n_redone = 0;
  if (!ticket) return CX_OK;  // _begin already ran the call to completion
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (h->shards) return shard_search_device_end(h, ticket, n_redone);
  std::unique_ptr<DeviceSearch> t((DeviceSearch*)ticket);
  CU(cudaSetDevice(h->device));
  Workspace* ws = t->lease.ws;
  const uint64_t before = h->fallbacks.load();
  cx_status st = run_search(h, ws, t->fh, t->sb, t->pl, false, 0.0f, nullptr, (uint32_t*)ws->hp, nullptr, nullptr,
                            RUN_FINISH);
  if (n_redone) *n_redone = h->fallbacks.load() - before;
  return st;
}

// ------------------------------------------------------------------------------
// The scan step of AutoLinker::run_cycle (linker/auto_linker.rs:215-264) for a batch of
// new nodes, restricted to what the similarity scan decides: search(emb, k) (k = 100 in the
// reference, :221), skip the node itself (:235-237), SimilarityLinkRule `score >= threshold`
// (linker/rules.rs:50), at most max_edges_per_node (:261).  Everything else in that loop
// (storage lookups, structural rules, dedupe against existing edges) stays on the host.
extern "C" cx_status cx_autolink_batch_device(cx_index* h, const float* d_embeddings, uint64_t B, uint64_t k,
                                              float threshold, uint32_t max_edges_per_node,
                                              const uint32_t* d_self_rows, uint32_t* d_scratch_rows,
                                              float* d_scratch_score, float* d_scratch_distance,
                                              uint32_t* d_scratch_n, uint32_t* d_out_rows, float* d_out_score,
                                              uint8_t* d_out_ids, uint32_t* d_out_n, void* stream) {
  cx::CallerDevice keep_callers_device;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (h->shards)
    return fail(CX_ERR_VALIDATION, "cx_autolink_batch_device: use cx_autolink_batch on a multi-device index");
  if (!d_scratch_rows || !d_scratch_score || !d_scratch_distance || !d_scratch_n || !d_out_rows || !d_out_score ||
      !d_out_n)
    return fail(CX_ERR_VALIDATION, "null device buffer");
  if (!B) return CX_OK;
  if (h->n_live == 0) {
    CU(cudaMemsetAsync(d_out_n, 0, B * 4, (cudaStream_t)stream));
    return CX_OK;
  }
  // Only the first max_edges + 1 neighbours can become links: the list is walked best first, the node itself
  // is skipped (at most one entry), the walk stops at the first score below the threshold or at the cap
  // (linker/auto_linker.rs:224-264).  Searching for min(k, max_edges + 1) is therefore exact and cheaper.
  uint64_t kk = k < h->n_rows ? k : h->n_rows;
  if (kk > (uint64_t)max_edges_per_node + 1) kk = (uint64_t)max_edges_per_node + 1;
  cx_status st = cx_search_batch_device(h, d_embeddings, B, kk, nullptr, d_scratch_rows, d_scratch_score,
                                        d_scratch_distance, nullptr, d_scratch_n, stream);
  if (st != CX_OK) return st;
  launch_autolink_filter(d_scratch_rows, d_scratch_score, d_scratch_n, d_self_rows, h->dIds, (uint32_t)B, (uint32_t)kk,
                         threshold, max_edges_per_node, d_out_rows, d_out_score, d_out_ids, d_out_n,
                         (cudaStream_t)stream);
  h->launches += 1;
  CU(cudaGetLastError());
  return CX_OK;
}

extern "C" cx_status cx_autolink_batch(cx_index* h, const uint8_t* new_ids, const float* embeddings, uint64_t B,
                                       uint32_t len, uint64_t k, float threshold, uint32_t max_edges_per_node,
                                       uint8_t* out_to_ids, float* out_score, uint32_t* out_n) {
  cx::CallerDevice keep_callers_device;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (!out_n || (B && (!embeddings || !out_to_ids || !out_score))) return fail(CX_ERR_VALIDATION, "null argument");
  if (len != h->dim)
    return fail(CX_ERR_VALIDATION, "Embedding dimension mismatch: expected %u, got %u", h->dim, len);
  for (uint64_t b = 0; b < B; ++b) out_n[b] = 0;
  if (h->shards)
    return shard_autolink_batch(h, new_ids, embeddings, B, k, threshold, max_edges_per_node, out_to_ids, out_score, out_n);
  if (!B || h->n_live == 0 || !max_edges_per_node || !k) return CX_OK;
  CU(cudaSetDevice(h->device));
  cx_status sst = settle(h);
  if (sst != CX_OK) return sst;
  const uint64_t kk = k < h->n_rows ? k : h->n_rows;
  const uint32_t me = max_edges_per_node;
  // everything this call stages lives in a leased workspace (no allocation per call once warm)
  WsLease outer(h);
  CU(outer.init());
  Workspace* wo = outer.ws;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_q = take(B * h->dim * 4), o_self = take(B * 4), o_rows = take(B * kk * 4), o_sc = take(B * kk * 4),
               o_di = take(B * kk * 4), o_n = take(B * 4), o_or = take(B * me * 4), o_os = take(B * me * 4),
               o_oi = take(B * me * 16), o_on = take(B * 4);
  // pinned mirror: [self rows | out score | out ids | out n]
  const size_t p_self = 0, p_os = align_up(p_self + B * 4, 256), p_oi = align_up(p_os + B * me * 4, 256),
               p_on = align_up(p_oi + B * me * 16, 256), p_total = p_on + B * 4;
  CU(wo->ensure_aux(off, p_total));
  char* d = (char*)wo->aux;
  char* hp = (char*)wo->aux_h;
  uint32_t* self = (uint32_t*)(hp + p_self);
  for (uint64_t b = 0; b < B; ++b) self[b] = 0xFFFFFFFFu;
  if (new_ids)
    for (uint64_t b = 0; b < B; ++b) {
      if (const uint32_t* row = h->id2row.find(load_id(new_ids + 16 * b))) self[b] = *row;
    }
  cudaStream_t s = wo->stream;
  CU(cudaMemcpyAsync(d + o_q, embeddings, B * h->dim * 4, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(d + o_self, self, B * 4, cudaMemcpyHostToDevice, s));
  cx_status st = cx_autolink_batch_device(h, (const float*)(d + o_q), B, kk, threshold, me, (const uint32_t*)(d + o_self),
                                          (uint32_t*)(d + o_rows), (float*)(d + o_sc), (float*)(d + o_di),
                                          (uint32_t*)(d + o_n), (uint32_t*)(d + o_or), (float*)(d + o_os),
                                          (uint8_t*)(d + o_oi), (uint32_t*)(d + o_on), s);
  if (st != CX_OK) return st;
  CU(cudaMemcpyAsync(hp + p_os, d + o_os, B * me * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(hp + p_oi, d + o_oi, B * me * 16, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(hp + p_on, d + o_on, B * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  memcpy(out_score, hp + p_os, B * me * 4);
  memcpy(out_to_ids, hp + p_oi, B * me * 16);
  memcpy(out_n, hp + p_on, B * 4);
  h->h2d += B * h->dim * 4;
  h->d2h += B * me * 20 + B * 4;
  return CX_OK;
}
