// cx_sharded.cu -- one index row-sharded over several GPUs of one box, driven from ONE process
// behind the same C ABI (cx_index_create_sharded): what `Arc<RwLock<V>>` in the reference's
// server (serve.rs:101) holds when V spans devices.
//
// The path shards naturally: every (query, row) score is independent and top-k /
// `>= threshold` lists are mergeable.  Each device owns a shard (a complete single-device index,
// cx_index.cu) plus the global insertion number of every row; inserts are spread so that the shards
// stay balanced.  A search fans out from one host thread per device (each issues its device's work
// on that device's streams), every device rescoring its own candidates exactly; the fixed-size
// local lists are then WRITTEN BY THE PRODUCING GPU straight into a buffer on devices[0] over
// NVLink (peer stores from the pack kernel; a peer copy where peer access is unavailable) and one
// kernel there merges them in the single-index order: score descending, NaN last, then global
// insertion order.  No NCCL: the exchange is 32 bytes per result, latency bound, and a store from
// the kernel that produced the value is the shortest path between two GPUs of an NVSwitch box.
//
// Results -- ids, score bits, order of equal scores -- are identical to those of a single-device
// index holding the same rows (tests/test_gpu_multidevice.py).
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <thread>

#include "cx_index.h"

namespace cx {

// ---- one host thread per shard ----------------------------------------------------------------
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  bool stop = false;
  void run(int device) {
    cudaSetDevice(device);
    for (;;) {
      std::function<void()> fn;
      {
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return stop || !q.empty(); });
        if (q.empty()) return;
        fn = std::move(q.front());
        q.pop_front();
      }
      fn();
    }
  }
  void post(std::function<void()> fn) {
    {
      std::lock_guard<std::mutex> g(mu);
      q.push_back(std::move(fn));
    }
    cv.notify_one();
  }
};

// Buffers of one call on the merge device (devices[0]): the gathered per-shard lists, the merged
// result and its pinned mirror.
struct MergeWs {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_done = nullptr;
  std::vector<cudaEvent_t> ev_shard;  // one per shard, recorded on the shard's stream after its pack
  void* d = nullptr;
  size_t d_bytes = 0;
  void* hp = nullptr;
  size_t h_bytes = 0;
  ~MergeWs() {
    if (d) cudaFree(d);
    if (hp) cudaFreeHost(hp);
    for (cudaEvent_t e : ev_shard)
      if (e) cudaEventDestroy(e);
    if (ev_done) cudaEventDestroy(ev_done);
    if (stream) cudaStreamDestroy(stream);
  }
};

struct ShardSet {
  std::vector<cx_index*> sh;
  std::vector<int> dev;
  std::vector<char> peer;  // shard s can store into the merge device's memory
  int merge_dev = 0;
  uint64_t next_seq = 0;
  IdMap id2shard;
  std::vector<std::unique_ptr<Worker>> workers;
  std::mutex mu;
  std::vector<MergeWs*> free_ws;
  uint32_t W() const { return (uint32_t)sh.size(); }
};

namespace {

struct Latch {
  std::mutex mu;
  std::condition_variable cv;
  uint32_t left;
  explicit Latch(uint32_t n) : left(n) {}
  void done() {
    std::lock_guard<std::mutex> g(mu);
    if (--left == 0) cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> g(mu);
    cv.wait(g, [&] { return left == 0; });
  }
};

// Run fn(s) for every shard on that shard's thread; returns the first failure (with its message
// re-raised on the calling thread).
cx_status for_shards(ShardSet* S, const std::function<cx_status(uint32_t)>& fn) {
  const uint32_t W = S->W();
  std::vector<cx_status> st(W, CX_OK);
  std::vector<std::string> msg(W);
  if (W == 1) {
    st[0] = fn(0);
    return st[0];
  }
  Latch latch(W);
  for (uint32_t s = 0; s < W; ++s)
    S->workers[s]->post([&, s] {
      st[s] = fn(s);
      if (st[s] != CX_OK) msg[s] = last_error();
      latch.done();
    });
  latch.wait();
  for (uint32_t s = 0; s < W; ++s)
    if (st[s] != CX_OK) return fail(st[s], "shard %u (device %d): %s", s, S->dev[s], msg[s].c_str());
  return CX_OK;
}

struct MergeLease {
  ShardSet* S;
  MergeWs* ws = nullptr;
  explicit MergeLease(ShardSet* S_) : S(S_) {
    std::lock_guard<std::mutex> g(S->mu);
    if (!S->free_ws.empty()) {
      ws = S->free_ws.back();
      S->free_ws.pop_back();
    }
  }
  cx_status init() {
    if (ws) return CX_OK;
    std::unique_ptr<MergeWs> w(new MergeWs());
    CU(cudaSetDevice(S->merge_dev));
    CU(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&w->ev_done, cudaEventDisableTiming));
    w->ev_shard.assign(S->W(), nullptr);
    for (uint32_t s = 0; s < S->W(); ++s) {
      CU(cudaSetDevice(S->dev[s]));
      CU(cudaEventCreateWithFlags(&w->ev_shard[s], cudaEventDisableTiming));
    }
    CU(cudaSetDevice(S->merge_dev));
    ws = w.release();
    return CX_OK;
  }
  cx_status ensure(size_t db, size_t hb) {
    CU(cudaSetDevice(S->merge_dev));
    if (db > ws->d_bytes) {
      if (ws->d) cudaFree(ws->d);
      ws->d = nullptr;
      ws->d_bytes = 0;
      CU(cudaMalloc(&ws->d, db + db / 4));
      ws->d_bytes = db + db / 4;
    }
    if (hb > ws->h_bytes) {
      if (ws->hp) cudaFreeHost(ws->hp);
      ws->hp = nullptr;
      ws->h_bytes = 0;
      CU(cudaMallocHost(&ws->hp, hb + hb / 4));
      ws->h_bytes = hb + hb / 4;
    }
    return CX_OK;
  }
  ~MergeLease() {
    if (!ws) return;
    std::lock_guard<std::mutex> g(S->mu);
    S->free_ws.push_back(ws);
  }
};

// ---- exchange kernels ---------------------------------------------------------------------------
// payload slot (32 B): w0 = score-order key << 32 | distance bits (0 = empty slot), w1 = global
// insertion number, w2/w3 = the 16 id bytes.  One trailer word per shard follows its B*k slots:
// how many of the shard's queries are still unverified.
constexpr uint32_t SLOT_WORDS = 4;

__global__ void shard_pack_kernel(const uint32_t* __restrict__ rows, const float* __restrict__ score,
                                  const float* __restrict__ dist, const uint32_t* __restrict__ n,
                                  const uint32_t* __restrict__ ok, uint32_t B, uint32_t k_local, uint32_t k,
                                  const uint64_t* __restrict__ seq, const uint8_t* __restrict__ ids,
                                  uint64_t* __restrict__ payload /* may live on another GPU */) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * k) return;
  const uint32_t b = i / k, j = i % k;
  if (ok && j == 0 && !ok[b]) atomicAdd(reinterpret_cast<unsigned long long*>(payload + (size_t)SLOT_WORDS * B * k), 1ull);
  uint64_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
  if (j < k_local && j < n[b]) {
    const size_t o = (size_t)b * k_local + j;
    const uint32_t r = rows[o];
    w0 = ((uint64_t)ord_from_score(score[o]) << 32) | (uint64_t)__float_as_uint(dist[o]);
    w1 = seq[r];
    const uint2* id = reinterpret_cast<const uint2*>(ids + (size_t)r * 16);
    const uint2 lo = id[0], hi = id[1];
    w2 = ((uint64_t)lo.y << 32) | lo.x;
    w3 = ((uint64_t)hi.y << 32) | hi.x;
  }
  // one 32-byte store per slot: a full sector on the wire
  ulonglong4 v;
  v.x = w0;
  v.y = w1;
  v.z = w2;
  v.w = w3;
  *reinterpret_cast<ulonglong4*>(payload + (size_t)SLOT_WORDS * i) = v;
}

// Does slot (x0, x1) come before (w0, w1) in the merged order?  (empty slots never do)
__device__ __forceinline__ bool slot_before(uint64_t x0, uint64_t x1, uint32_t ord, uint64_t sq) {
  if (x0 == 0) return false;
  const uint32_t xo = (uint32_t)(x0 >> 32);
  return xo > ord || (xo == ord && x1 < sq);
}

// Every shard's list is already in merged order (its valid slots form a prefix), so the rank of a
// slot is its own position plus, for every other shard, the number of that shard's slots that come
// before it: one binary search per (slot, other shard).  No shared memory, any k.
__global__ void shard_merge_kernel(const uint64_t* __restrict__ gathered, uint32_t W, uint32_t B, uint32_t k,
                                   uint32_t k_out, uint32_t* __restrict__ out_rows, float* __restrict__ out_score,
                                   float* __restrict__ out_dist, uint8_t* __restrict__ out_ids,
                                   uint64_t* __restrict__ out_seq, uint32_t* __restrict__ out_n,
                                   uint64_t* __restrict__ out_unverified) {
  const uint32_t b = blockIdx.x;
  const size_t n_slot_words = (size_t)SLOT_WORDS * B * k;
  const size_t stride = n_slot_words + SLOT_WORDS;  // words per shard: slots + trailer, 32 B aligned
  if (b == 0 && threadIdx.x == 0 && out_unverified) {
    uint64_t t = 0;
    for (uint32_t w = 0; w < W; ++w) t += gathered[(size_t)w * stride + n_slot_words];
    *out_unverified = t;
  }
  uint32_t valid = 0;
  for (uint32_t i = threadIdx.x; i < W * k; i += blockDim.x) {
    const uint32_t w = i / k, j = i % k;
    const uint64_t* mine = gathered + (size_t)w * stride + (size_t)SLOT_WORDS * ((size_t)b * k + j);
    const ulonglong4 v = *reinterpret_cast<const ulonglong4*>(mine);
    if (v.x == 0) continue;
    ++valid;
    const uint32_t ord = (uint32_t)(v.x >> 32);
    uint32_t rank = j;
    for (uint32_t x = 0; x < W; ++x) {
      if (x == w) continue;
      const uint64_t* L = gathered + (size_t)x * stride + (size_t)SLOT_WORDS * (size_t)b * k;
      uint32_t lo = 0, hi = k;  // first slot of shard x that does NOT come before mine
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (slot_before(L[(size_t)SLOT_WORDS * mid], L[(size_t)SLOT_WORDS * mid + 1], ord, v.y)) lo = mid + 1;
        else hi = mid;
      }
      rank += lo;
    }
    if (rank < k_out) {
      const size_t o = (size_t)b * k_out + rank;
      const float d = __uint_as_float((uint32_t)v.x);
      if (out_rows) out_rows[o] = (uint32_t)v.y;
      if (out_seq) out_seq[o] = v.y;
      out_dist[o] = d;
      out_score[o] = ref_score_from_distance(d);  // the reference's own two operations (index.rs:254-256)
      if (out_ids) {
        uint64_t* id = reinterpret_cast<uint64_t*>(out_ids + o * 16);
        id[0] = v.z;
        id[1] = v.w;
      }
    }
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  __shared__ uint32_t s_valid;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&s_valid, valid);
  __syncthreads();
  if (threadIdx.x == 0) out_n[b] = s_valid < k_out ? s_valid : k_out;
}

// auto-link candidate post-pass on merged lists (linker/auto_linker.rs:224-264, rules.rs:42-62): walk the
// node's neighbours best first, skip the node itself (by global insertion number), keep
// score >= threshold, stop at max_edges.
__global__ void shard_autolink_kernel(const uint64_t* __restrict__ seq, const float* __restrict__ score,
                                      const uint8_t* __restrict__ ids, const uint32_t* __restrict__ n,
                                      const uint64_t* __restrict__ self_seq, uint32_t B, uint32_t k, float threshold,
                                      uint32_t max_edges, float* __restrict__ out_score, uint8_t* __restrict__ out_ids,
                                      uint32_t* __restrict__ out_n) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint64_t self = self_seq ? self_seq[b] : ~0ull;
  const uint32_t nb = n[b] < k ? n[b] : k;
  uint32_t m = 0;
  for (uint32_t j = 0; j < nb && m < max_edges; ++j) {
    const size_t i = (size_t)b * k + j;
    if (seq[i] == self) continue;
    const float s = score[i];
    if (!(s >= threshold)) continue;  // false for NaN
    const size_t o = (size_t)b * max_edges + m;
    out_score[o] = s;
    *reinterpret_cast<uint4*>(out_ids + o * 16) = *reinterpret_cast<const uint4*>(ids + i * 16);
    ++m;
  }
  out_n[b] = m;
}

// dedup: merged per-node partner lists -> dense (node index in block, partner id, score) records
__global__ void shard_pair_write_kernel(const uint8_t* __restrict__ ids, const float* __restrict__ score,
                                        const uint32_t* __restrict__ n, const uint32_t* __restrict__ off, uint32_t B,
                                        uint32_t kd, uint32_t* __restrict__ rec /* 6 words per pair */,
                                        uint32_t limit) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B) return;
  const uint32_t m = n[w], base = off[w];
  for (uint32_t j = lane; j < m; j += 32) {
    const uint32_t pos = base + j;
    if (pos >= limit) break;
    const uint32_t* id = reinterpret_cast<const uint32_t*>(ids + ((size_t)w * kd + j) * 16);
    uint32_t* o = rec + 6 * (size_t)pos;
    o[0] = w;
    o[1] = __float_as_uint(score[(size_t)w * kd + j]);
    o[2] = id[0];
    o[3] = id[1];
    o[4] = id[2];
    o[5] = id[3];
  }
}

// slots + the trailer word, padded so that every shard's region starts 32 B aligned
size_t payload_words(uint64_t B, uint64_t k) { return (size_t)SLOT_WORDS * B * k + SLOT_WORDS; }

// Per-shard part of a fan-out: device buffers of the call on the shard's device, carved from the
// aux block of a workspace leased from the shard.
struct ShardCall {
  std::unique_ptr<WsLease> outer;
  float* dQ = nullptr;
  uint32_t *rows = nullptr, *n = nullptr, *tot = nullptr;
  float *score = nullptr, *dist = nullptr;
  uint64_t* payload = nullptr;  // local staging when the shard cannot store into the merge device
  uint64_t* q_seq = nullptr;
  void* ticket = nullptr;
  uint32_t kd = 0;
  uint64_t redone = 0;
};

cx_status shard_call_init(cx_index* c, ShardCall* sc, uint64_t B, uint32_t qlen, uint64_t k, bool need_payload,
                          bool need_seq) {
  CU(cudaSetDevice(c->device));
  sc->outer.reset(new WsLease(c));
  CU(sc->outer->init());
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_q = take(B * qlen * 4), o_r = take(B * k * 4), o_s = take(B * k * 4), o_d = take(B * k * 4),
               o_n = take(B * 4), o_t = take(B * 4), o_p = take(need_payload ? payload_words(B, k) * 8 : 0),
               o_qs = take(need_seq ? B * 8 : 0);
  CU(sc->outer->ws->ensure_aux(off, 256));
  char* d = (char*)sc->outer->ws->aux;
  sc->dQ = (float*)(d + o_q);
  sc->rows = (uint32_t*)(d + o_r);
  sc->score = (float*)(d + o_s);
  sc->dist = (float*)(d + o_d);
  sc->n = (uint32_t*)(d + o_n);
  sc->tot = (uint32_t*)(d + o_t);
  sc->payload = need_payload ? (uint64_t*)(d + o_p) : nullptr;
  sc->q_seq = need_seq ? (uint64_t*)(d + o_qs) : nullptr;
  return CX_OK;
}

// pack this shard's local lists into its region of the gathered buffer on the merge device
cx_status shard_pack(ShardSet* S, uint32_t s, ShardCall& sc, MergeWs* mw, uint64_t* gathered, uint64_t B, uint64_t k,
                     const uint32_t* d_ok) {
  cx_index* c = S->sh[s];
  cudaStream_t st = sc.outer->ws->stream;
  const size_t words = payload_words(B, k);
  uint64_t* dst = gathered + (size_t)s * words;
  uint64_t* target = S->peer[s] ? dst : sc.payload;
  CU(cudaMemsetAsync(target + words - SLOT_WORDS, 0, 8 * SLOT_WORDS, st));  // trailer
  const uint64_t total = B * k;
  if (c->n_rows == 0) {
    CU(cudaMemsetAsync(target, 0, words * 8, st));
  } else {
    shard_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sc.rows, sc.score, sc.dist, sc.n, d_ok, (uint32_t)B,
                                                                       sc.kd, (uint32_t)k, c->dSeq, c->dIds, target);
    CU(cudaGetLastError());
    c->launches += 1;
  }
  if (!S->peer[s]) CU(cudaMemcpyPeerAsync(dst, S->merge_dev, sc.payload, c->device, words * 8, st));
  CU(cudaEventRecord(mw->ev_shard[s], st));
  return CX_OK;
}

struct MergedBufs {
  uint64_t* gathered;
  uint32_t* rows;
  float *score, *dist;
  uint8_t* ids;
  uint64_t* seq;
  uint32_t* n;
  uint64_t* unverified;
  size_t dev_bytes;
  // pinned mirror
  size_t h_score, h_dist, h_ids, h_n, h_unv, h_bytes;
};

MergedBufs carve_merged(void* base, uint32_t W, uint64_t B, uint64_t k, uint64_t k_out) {
  MergedBufs m;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  char* d = (char*)base;
  const size_t o_g = take((size_t)W * payload_words(B, k) * 8), o_r = take(B * k_out * 4), o_s = take(B * k_out * 4),
               o_d = take(B * k_out * 4), o_i = take(B * k_out * 16), o_q = take(B * k_out * 8), o_n = take(B * 4),
               o_u = take(8);
  m.gathered = (uint64_t*)(d + o_g);
  m.rows = (uint32_t*)(d + o_r);
  m.score = (float*)(d + o_s);
  m.dist = (float*)(d + o_d);
  m.ids = (uint8_t*)(d + o_i);
  m.seq = (uint64_t*)(d + o_q);
  m.n = (uint32_t*)(d + o_n);
  m.unverified = (uint64_t*)(d + o_u);
  m.dev_bytes = off;
  size_t h = 0;
  auto htake = [&](size_t bytes) { size_t o = h; h = align_up(h + bytes, 256); return o; };
  m.h_score = htake(B * k_out * 4);
  m.h_dist = htake(B * k_out * 4);
  m.h_ids = htake(B * k_out * 16);
  m.h_n = htake(B * 4);
  m.h_unv = htake(8);
  m.h_bytes = h;
  return m;
}

cx_status launch_merge(ShardSet* S, MergeWs* mw, const MergedBufs& m, uint64_t B, uint64_t k, uint64_t k_out,
                       uint32_t* rows, float* score, float* dist, uint8_t* ids, uint64_t* seq, uint32_t* n) {
  CU(cudaSetDevice(S->merge_dev));
  for (uint32_t s = 0; s < S->W(); ++s) CU(cudaStreamWaitEvent(mw->stream, mw->ev_shard[s], 0));
  const uint32_t threads = (uint32_t)std::min<uint64_t>(256, align_up(S->W() * k, 32));
  shard_merge_kernel<<<(unsigned)B, threads, 0, mw->stream>>>(m.gathered, S->W(), (uint32_t)B, (uint32_t)k, (uint32_t)k_out,
                                                              rows, score, dist, ids, seq, n, m.unverified);
  CU(cudaGetLastError());
  S->sh[0]->launches += 1;
  return CX_OK;
}

}  // namespace

// ---- lifecycle -------------------------------------------------------------------------------------
void shard_destroy(cx_index* h) {
  ShardSet* S = h->shards;
  if (!S) return;
  for (auto& w : S->workers) {
    {
      std::lock_guard<std::mutex> g(w->mu);
      w->stop = true;
    }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
  }
  cudaSetDevice(S->merge_dev);
  for (MergeWs* w : S->free_ws) delete w;
  for (cx_index* c : S->sh) cx_index_destroy(c);
  delete S;
  h->shards = nullptr;
}

uint64_t shard_len(const cx_index* h) {
  uint64_t n = 0;
  for (const cx_index* c : h->shards->sh) n += c->n_live;
  return n;
}

cx_status shard_reserve(cx_index* h, uint64_t n_rows) {
  ShardSet* S = h->shards;
  const uint64_t per = (n_rows + S->W() - 1) / S->W();
  return for_shards(S, [&](uint32_t s) { return cx_reserve(S->sh[s], per); });
}

// New rows are dealt out in contiguous blocks of the batch, filling the emptiest shards first, so that
// the shards stay balanced whether rows arrive one by one, in streaming batches or as one bulk load.
// Ids that already exist are overwritten where they live.  The global insertion number of a row is
// the tie order of equal scores (the single index's row order).
cx_status shard_insert(cx_index* h, const uint8_t* ids, const float* rows, uint64_t n, uint32_t len, bool on_device) {
  ShardSet* S = h->shards;
  if (len != h->dim)
    return fail(CX_ERR_VALIDATION, "Embedding dimension mismatch: expected %u, got %u", h->dim, len);
  if (!n) return CX_OK;
  if (!ids || !rows) return fail(CX_ERR_VALIDATION, "null ids/rows");
  const uint32_t W = S->W();
  // where does every row go?
  std::vector<uint16_t> where(n);
  std::vector<uint64_t> fresh_pos;  // batch positions of first occurrences of new ids
  {
    std::unordered_map<Id128, uint64_t, Id128Hash> seen;  // new id -> index into fresh_pos
    std::vector<uint64_t> dup_of(n, ~0ull);
    for (uint64_t i = 0; i < n; ++i) {
      const Id128 key = load_id(ids + 16 * i);
      if (const uint32_t* sh_of = S->id2shard.find(key)) {
        where[i] = (uint16_t)*sh_of;
        continue;
      }
      auto f = seen.find(key);
      if (f != seen.end()) {
        dup_of[i] = f->second;
        continue;
      }
      seen.emplace(key, fresh_pos.size());
      fresh_pos.push_back(i);
      where[i] = 0xFFFF;
    }
    // deal the new rows: shard s takes what it lacks to the common level, in shard order
    const uint64_t m = fresh_pos.size();
    uint64_t total = m;
    for (cx_index* c : S->sh) total += c->n_rows;
    const uint64_t level = (total + W - 1) / W;
    std::vector<uint16_t> fresh_shard(m);
    uint64_t at = 0;
    for (uint32_t s = 0; s < W && at < m; ++s) {
      const uint64_t have = S->sh[s]->n_rows;
      uint64_t take = have < level ? level - have : 0;
      if (take > m - at) take = m - at;
      for (uint64_t j = 0; j < take; ++j) fresh_shard[at + j] = (uint16_t)s;
      at += take;
    }
    for (; at < m; ++at) fresh_shard[at] = (uint16_t)(W - 1);
    for (uint64_t j = 0; j < m; ++j) where[fresh_pos[j]] = fresh_shard[j];
    for (uint64_t i = 0; i < n; ++i)
      if (dup_of[i] != ~0ull) where[i] = fresh_shard[dup_of[i]];
  }
  // per shard: the batch positions it receives, in batch order; contiguous runs go down as one call
  const uint64_t seq0 = S->next_seq;
  std::vector<uint64_t> seq(n);
  for (uint64_t i = 0; i < n; ++i) seq[i] = seq0 + i;
  std::vector<std::vector<std::pair<uint64_t, uint64_t>>> runs(W);  // (first, count)
  for (uint64_t i = 0; i < n;) {
    uint64_t j = i + 1;
    while (j < n && where[j] == where[i]) ++j;
    runs[where[i]].push_back({i, j - i});
    i = j;
  }
  const int src_dev = S->merge_dev;
  cx_status st = for_shards(S, [&](uint32_t s) -> cx_status {
    cx_index* c = S->sh[s];
    for (auto& r : runs[s]) {
      const float* src = rows + (size_t)r.first * len;
      if (on_device && c->device != src_dev) {
        // rows live on devices[0]: bring this shard's part over NVLink, then insert from local memory
        CU(cudaSetDevice(c->device));
        WsLease tmp(c);
        CU(tmp.init());
        CU(tmp.ws->ensure_aux(r.second * len * 4, 256));
        CU(cudaMemcpyPeerAsync(tmp.ws->aux, c->device, src, src_dev, r.second * len * 4, tmp.ws->stream));
        CU(cudaStreamSynchronize(tmp.ws->stream));
        cx_status e = index_insert(c, ids + 16 * r.first, (const float*)tmp.ws->aux, r.second, len, true, seq.data() + r.first);
        if (e != CX_OK) return e;
      } else {
        cx_status e = index_insert(c, ids + 16 * r.first, src, r.second, len, on_device, seq.data() + r.first);
        if (e != CX_OK) return e;
      }
    }
    return CX_OK;
  });
  // record what actually arrived (after a failure on one device the other devices' rows are in)
  for (uint64_t i = 0; i < n; ++i) {
    const Id128 key = load_id(ids + 16 * i);
    if (st != CX_OK && !S->sh[where[i]]->id2row.count(key)) continue;
    if (S->id2shard.emplace(key, where[i]))
      for (uint32_t s = 0; s < W; ++s)  // metadata that waited for this id was applied by the receiving shard
        if (s != where[i] && !S->sh[s]->orphan_meta.empty()) S->sh[s]->orphan_meta.erase(key);
  }
  S->next_seq = seq0 + n;
  return st;
}

cx_status shard_remove(cx_index* h, const uint8_t id[16]) {
  ShardSet* S = h->shards;
  const Id128 key = load_id(id);
  const uint32_t* sh_of = S->id2shard.find(key);
  if (!sh_of) {
    for (cx_index* c : S->sh) c->orphan_meta.erase(key);
    return CX_OK;
  }
  cx_status st = index_remove(S->sh[*sh_of], id);
  if (st == CX_OK) S->id2shard.erase(key);
  return st;
}

cx_status shard_set_metadata(cx_index* h, const uint8_t id[16], const char* kind, const char* agent) {
  ShardSet* S = h->shards;
  if (const uint32_t* sh_of = S->id2shard.find(load_id(id))) return index_set_metadata(S->sh[*sh_of], id, kind, agent);
  // metadata for an id that has no vector (yet): every shard remembers it, whichever receives the row
  // later applies it.  Kind / agent strings are interned per shard, filters are built per shard.
  for (cx_index* c : S->sh) {
    cx_status st = index_set_metadata(c, id, kind, agent);
    if (st != CX_OK) return st;
  }
  return CX_OK;
}

cx_status shard_rebuild(cx_index* h) {
  ShardSet* S = h->shards;
  return for_shards(S, [&](uint32_t s) { return index_rebuild(S->sh[s]); });
}

cx_status shard_set_option(cx_index* h, const char* key, int64_t value) {
  for (cx_index* c : h->shards->sh) {
    cx_status st = index_set_option(c, key, value);
    if (st != CX_OK) return st;
  }
  return CX_OK;
}

cx_status shard_stats(cx_index* h, cx_stats* out) {
  for (cx_index* c : h->shards->sh) {
    cx_status st = settle(c);
    if (st != CX_OK) return st;
    index_add_stats(c, out);
  }
  return CX_OK;
}

// save: the single-index file (index.rs:437-472), rows in global insertion order
cx_status shard_save(cx_index* h, const char* path) {
  ShardSet* S = h->shards;
  FILE* fp = fopen(path, "wb");
  if (!fp) return fail(CX_ERR_IO, "Failed to write index file: %s", path);
  auto w64 = [&](uint64_t v) { return fwrite(&v, 8, 1, fp) == 1; };
  bool ok = w64(shard_len(h));
  // W-way merge of the shards' rows by insertion number; runs of consecutive rows of one shard are
  // written together
  const uint32_t W = S->W();
  std::vector<uint64_t> at(W, 0);
  while (ok) {
    uint32_t best = W;
    for (uint32_t s = 0; s < W; ++s)
      if (at[s] < S->sh[s]->n_rows && (best == W || S->sh[s]->h_seq[at[s]] < S->sh[best]->h_seq[at[best]])) best = s;
    if (best == W) break;
    uint64_t limit = ~0ull;  // the run ends before the next row of any other shard
    for (uint32_t s = 0; s < W; ++s)
      if (s != best && at[s] < S->sh[s]->n_rows && S->sh[s]->h_seq[at[s]] < limit) limit = S->sh[s]->h_seq[at[s]];
    cx_index* c = S->sh[best];
    uint64_t e = at[best] + 1;
    while (e < c->n_rows && c->h_seq[e] < limit) ++e;
    if (index_save_vectors(c, fp, at[best], e) != CX_OK) ok = false;
    at[best] = e;
  }
  uint64_t n_meta = 0;
  // metadata without a vector is remembered by every shard: count it once (shard 0's view)
  for (uint32_t s = 0; s < W; ++s) n_meta += index_meta_count(S->sh[s]) - (s ? S->sh[s]->orphan_meta.size() : 0);
  ok = ok && w64(n_meta);
  for (uint32_t s = 0; s < W && ok; ++s) {
    if (s == 0) {
      ok = index_save_meta(S->sh[0], fp);
    } else {
      auto keep = std::move(S->sh[s]->orphan_meta);  // written with shard 0
      S->sh[s]->orphan_meta.clear();
      ok = index_save_meta(S->sh[s], fp);
      S->sh[s]->orphan_meta = std::move(keep);
    }
  }
  ok = ok && w64(h->dim);
  ok = (fclose(fp) == 0) && ok;
  if (!ok) return fail(CX_ERR_IO, "Failed to write index file: %s", path);
  return CX_OK;
}

// ---- searches -----------------------------------------------------------------------------------------
namespace {

struct FanOut {
  ShardSet* S;
  MergeLease ml;
  std::vector<ShardCall> calls;
  MergedBufs m;
  uint64_t B = 0, k = 0;
  explicit FanOut(ShardSet* S_) : S(S_), ml(S_), calls(S_->W()) {}
};

// top-k over all shards with queries in host memory (q_host) or on the merge device (q_dev).
// Enqueues everything up to the merge; `after` (optional) runs on the merge stream behind it.
cx_status fan_topk_begin(cx_index* h, FanOut& f, const float* q_host, const float* q_dev, cudaStream_t q_ready_on,
                         uint64_t B, uint32_t qlen, uint64_t k, const cx_filter* filter) {
  ShardSet* S = f.S;
  f.B = B;
  f.k = k;
  cx_status st = f.ml.init();
  if (st != CX_OK) return st;
  f.m = carve_merged(nullptr, S->W(), B, k, k);
  st = f.ml.ensure(f.m.dev_bytes, f.m.h_bytes);
  if (st != CX_OK) return st;
  f.m = carve_merged(f.ml.ws->d, S->W(), B, k, k);
  MergeWs* mw = f.ml.ws;
  cudaEvent_t q_ready = nullptr;
  if (q_dev) {  // the shards' copies of the queries must wait for whatever produced them
    CU(cudaSetDevice(S->merge_dev));
    CU(cudaEventRecord(mw->ev_done, q_ready_on));
    q_ready = mw->ev_done;
  }
  return for_shards(S, [&](uint32_t s) -> cx_status {
    cx_index* c = S->sh[s];
    ShardCall& sc = f.calls[s];
    cx_status e = settle(c);
    if (e != CX_OK) return e;
    sc.kd = (uint32_t)std::min<uint64_t>(k, c->n_rows);
    e = shard_call_init(c, &sc, B, qlen, k, !S->peer[s], false);
    if (e != CX_OK) return e;
    cudaStream_t stx = sc.outer->ws->stream;
    sc.ticket = nullptr;
    if (c->n_live == 0 || sc.kd == 0) {
      CU(cudaMemsetAsync(sc.n, 0, B * 4, stx));
    } else {
      if (q_dev) {
        CU(cudaStreamWaitEvent(stx, q_ready, 0));
        CU(cudaMemcpyPeerAsync(sc.dQ, c->device, q_dev, S->merge_dev, B * qlen * 4, stx));
      } else {
        CU(cudaMemcpyAsync(sc.dQ, q_host, B * qlen * 4, cudaMemcpyHostToDevice, stx));
        c->h2d += B * qlen * 4;
      }
      e = index_search_device(c, sc.dQ, qlen, B, sc.kd, filter, sc.rows, sc.score, sc.dist, nullptr, sc.n, stx, &sc.ticket);
      if (e != CX_OK) return e;
    }
    const uint32_t* d_ok = nullptr;
    cx_search_ticket_ok(sc.ticket, &d_ok);
    return shard_pack(S, s, sc, mw, f.m.gathered, B, k, d_ok);
  });
}

// finish every shard's search (retries of unverified queries happen here); if any shard changed its
// lists, pack those again.  Returns the number of queries redone over all shards.
cx_status fan_topk_end(FanOut& f, uint64_t* redone_total) {
  ShardSet* S = f.S;
  MergeWs* mw = f.ml.ws;
  cx_status st = for_shards(S, [&](uint32_t s) -> cx_status {
    ShardCall& sc = f.calls[s];
    sc.redone = 0;
    if (!sc.ticket) return CX_OK;
    cx_status e = cx_search_batch_device_end(S->sh[s], sc.ticket, &sc.redone);
    sc.ticket = nullptr;
    if (e != CX_OK) return e;
    if (sc.redone) return shard_pack(S, s, sc, mw, f.m.gathered, f.B, f.k, nullptr);
    return CX_OK;
  });
  uint64_t t = 0;
  for (auto& sc : f.calls) t += sc.redone;
  if (redone_total) *redone_total = t;
  return st;
}

}  // namespace

cx_status shard_search_host(cx_index* h, const float* queries, uint64_t B, uint32_t qlen, uint64_t k,
                            const cx_filter* filter, bool threshold_mode, float threshold, uint8_t* out_ids,
                            float* out_score, float* out_dist, uint64_t* out_n, uint64_t* out_total) {
  ShardSet* S = h->shards;
  const uint64_t n_live = shard_len(h);
  if (n_live == 0 || B == 0 || k == 0) return CX_OK;  // index.rs:331
  uint64_t n_rows = 0;
  for (cx_index* c : S->sh) n_rows += c->n_rows;
  const uint64_t kk = k < n_rows ? k : n_rows;  // slots per query
  FanOut f(S);
  MergeWs* mw = nullptr;
  cx_status st;
  std::vector<std::vector<uint64_t>> totals;
  if (!threshold_mode) {
    st = fan_topk_begin(h, f, queries, nullptr, nullptr, B, qlen, kk, filter);
    if (st != CX_OK) return st;
    mw = f.ml.ws;
    st = launch_merge(S, mw, f.m, B, kk, kk, nullptr, f.m.score, f.m.dist, f.m.ids, nullptr, f.m.n);
    if (st != CX_OK) return st;
    uint64_t redone = 0;
    st = fan_topk_end(f, &redone);  // waits for every shard's scan
    if (st != CX_OK) return st;
    if (redone) {
      st = launch_merge(S, mw, f.m, B, kk, kk, nullptr, f.m.score, f.m.dist, f.m.ids, nullptr, f.m.n);
      if (st != CX_OK) return st;
    }
  } else {
    // every shard: all of its rows with score >= threshold, best first (at most kk of them) + how many qualify
    f.B = B;
    f.k = kk;
    st = f.ml.init();
    if (st != CX_OK) return st;
    f.m = carve_merged(nullptr, S->W(), B, kk, kk);
    st = f.ml.ensure(f.m.dev_bytes, f.m.h_bytes);
    if (st != CX_OK) return st;
    f.m = carve_merged(f.ml.ws->d, S->W(), B, kk, kk);
    mw = f.ml.ws;
    totals.assign(S->W(), std::vector<uint64_t>(B, 0));
    st = for_shards(S, [&](uint32_t s) -> cx_status {
      cx_index* c = S->sh[s];
      ShardCall& sc = f.calls[s];
      sc.kd = (uint32_t)kk;
      cx_status e = shard_call_init(c, &sc, B, qlen, kk, !S->peer[s], false);
      if (e != CX_OK) return e;
      cudaStream_t stx = sc.outer->ws->stream;
      CU(cudaMemcpyAsync(sc.dQ, queries, B * qlen * 4, cudaMemcpyHostToDevice, stx));
      CU(cudaStreamSynchronize(stx));
      c->h2d += B * qlen * 4;
      e = index_threshold_device(c, sc.dQ, qlen, qlen, B, threshold, filter, kk, sc.rows, sc.score, sc.dist, nullptr,
                                 sc.n, nullptr, totals[s].data(), nullptr);
      if (e != CX_OK) return e;
      return shard_pack(S, s, sc, mw, f.m.gathered, B, kk, nullptr);
    });
    if (st != CX_OK) return st;
    st = launch_merge(S, mw, f.m, B, kk, kk, nullptr, f.m.score, f.m.dist, f.m.ids, nullptr, f.m.n);
    if (st != CX_OK) return st;
  }
  // one copy of the merged block back
  CU(cudaSetDevice(S->merge_dev));
  char* hp = (char*)mw->hp;
  CU(cudaMemcpyAsync(hp + f.m.h_score, f.m.score, B * kk * 4, cudaMemcpyDeviceToHost, mw->stream));
  CU(cudaMemcpyAsync(hp + f.m.h_dist, f.m.dist, B * kk * 4, cudaMemcpyDeviceToHost, mw->stream));
  CU(cudaMemcpyAsync(hp + f.m.h_ids, f.m.ids, B * kk * 16, cudaMemcpyDeviceToHost, mw->stream));
  CU(cudaMemcpyAsync(hp + f.m.h_n, f.m.n, B * 4, cudaMemcpyDeviceToHost, mw->stream));
  CU(cudaStreamSynchronize(mw->stream));
  S->sh[0]->d2h += B * kk * 24 + B * 4;
  const uint32_t* h_n = (const uint32_t*)(hp + f.m.h_n);
  for (uint64_t b = 0; b < B; ++b) {
    const uint32_t n = h_n[b];
    out_n[b] = n;
    if (out_total) {
      uint64_t t = n;
      if (threshold_mode) {
        t = 0;
        for (uint32_t s = 0; s < S->W(); ++s) t += totals[s][b];
      }
      out_total[b] = t;
    }
    if (out_score) memcpy(out_score + b * k, hp + f.m.h_score + b * kk * 4, (size_t)n * 4);
    if (out_dist) memcpy(out_dist + b * k, hp + f.m.h_dist + b * kk * 4, (size_t)n * 4);
    if (out_ids) memcpy(out_ids + b * k * 16, hp + f.m.h_ids + b * kk * 16, (size_t)n * 16);
  }
  return CX_OK;
}

// Device-resident form: queries [B][dim] and every output buffer live on devices[0].  d_out_rows
// receives the low 32 bits of each result's global insertion number.
struct ShardTicket {
  FanOut f;
  cudaStream_t user;
  uint32_t* rows;
  float *score, *dist;
  uint8_t* ids;
  uint32_t* n;
  explicit ShardTicket(ShardSet* S) : f(S) {}
};

cx_status shard_search_device(cx_index* h, const float* d_queries, uint64_t B, uint64_t k, const cx_filter* filter,
                              uint32_t* d_out_rows, float* d_out_score, float* d_out_distance, uint8_t* d_out_ids,
                              uint32_t* d_out_n, void* stream, void** ticket) {
  ShardSet* S = h->shards;
  if (ticket) *ticket = nullptr;
  if (!d_queries || !d_out_rows || !d_out_score || !d_out_distance || !d_out_n)
    return fail(CX_ERR_VALIDATION, "null device buffer");
  if (B == 0) return CX_OK;
  cudaStream_t user = (cudaStream_t)stream;
  CU(cudaSetDevice(S->merge_dev));
  if (shard_len(h) == 0 || k == 0) {
    CU(cudaMemsetAsync(d_out_n, 0, B * 4, user));
    return CX_OK;
  }
  uint64_t n_rows = 0;
  for (cx_index* c : S->sh) n_rows += c->n_rows;
  if (k > n_rows) return fail(CX_ERR_VALIDATION, "device search needs k <= rows in the index");
  std::unique_ptr<ShardTicket> t(new ShardTicket(S));
  t->user = user;
  t->rows = d_out_rows;
  t->score = d_out_score;
  t->dist = d_out_distance;
  t->ids = d_out_ids;
  t->n = d_out_n;
  cx_status st = fan_topk_begin(h, t->f, nullptr, d_queries, user, B, h->dim, k, filter);
  if (st != CX_OK) return st;
  MergeWs* mw = t->f.ml.ws;
  st = launch_merge(S, mw, t->f.m, B, k, k, d_out_rows, d_out_score, d_out_distance, d_out_ids, nullptr, d_out_n);
  if (st != CX_OK) return st;
  CU(cudaEventRecord(mw->ev_done, mw->stream));
  CU(cudaStreamWaitEvent(user, mw->ev_done, 0));
  if (ticket) {
    *ticket = t.release();
    return CX_OK;
  }
  return shard_search_device_end(h, t.release(), nullptr);
}

cx_status shard_search_device_end(cx_index* h, void* ticket, uint64_t* n_redone) {
  ShardSet* S = h->shards;
  std::unique_ptr<ShardTicket> t((ShardTicket*)ticket);
  uint64_t redone = 0;
  cx_status st = fan_topk_end(t->f, &redone);
  if (st != CX_OK) return st;
  MergeWs* mw = t->f.ml.ws;
  CU(cudaSetDevice(S->merge_dev));
  if (redone) {
    st = launch_merge(S, mw, t->f.m, t->f.B, t->f.k, t->f.k, t->rows, t->score, t->dist, t->ids, nullptr, t->n);
    if (st != CX_OK) return st;
  }
  CU(cudaStreamSynchronize(mw->stream));
  if (n_redone) *n_redone = redone;
  return CX_OK;
}

// AutoLinker::run_cycle's scan step (linker/auto_linker.rs:215-264) over the sharded corpus: sharded
// search(k), then ONE kernel on the merged lists applies skip-self / threshold / per-node cap.
cx_status shard_autolink_batch(cx_index* h, const uint8_t* new_ids, const float* embeddings, uint64_t B, uint64_t k,
                               float threshold, uint32_t max_edges_per_node, uint8_t* out_to_ids, float* out_score,
                               uint32_t* out_n) {
  ShardSet* S = h->shards;
  if (!B || shard_len(h) == 0 || !max_edges_per_node || !k) return CX_OK;
  uint64_t n_rows = 0;
  for (cx_index* c : S->sh) n_rows += c->n_rows;
  const uint32_t me = max_edges_per_node;
  uint64_t kk = k < n_rows ? k : n_rows;
  if (kk > (uint64_t)me + 1) kk = (uint64_t)me + 1;  // only the first me + 1 neighbours can become links (exact)
  FanOut f(S);
  cx_status st = fan_topk_begin(h, f, embeddings, nullptr, nullptr, B, h->dim, kk, nullptr);
  if (st != CX_OK) return st;
  MergeWs* mw = f.ml.ws;
  st = launch_merge(S, mw, f.m, B, kk, kk, nullptr, f.m.score, f.m.dist, f.m.ids, f.m.seq, f.m.n);
  if (st != CX_OK) return st;
  uint64_t redone = 0;
  st = fan_topk_end(f, &redone);
  if (st != CX_OK) return st;
  if (redone) {
    st = launch_merge(S, mw, f.m, B, kk, kk, nullptr, f.m.score, f.m.dist, f.m.ids, f.m.seq, f.m.n);
    if (st != CX_OK) return st;
  }
  // post-pass buffers on the merge device: [self seq | out score | out ids | out n], same layout pinned
  CU(cudaSetDevice(S->merge_dev));
  WsLease aux(S->sh[0]);  // shard 0 lives on the merge device
  CU(aux.init());
  const size_t o_self = 0, o_os = align_up(o_self + B * 8, 256), o_oi = align_up(o_os + B * me * 4, 256),
               o_on = align_up(o_oi + B * me * 16, 256), total = o_on + B * 4;
  CU(aux.ws->ensure_aux(total, total));
  char* d = (char*)aux.ws->aux;
  char* hp = (char*)aux.ws->aux_h;
  uint64_t* self = (uint64_t*)(hp + o_self);
  for (uint64_t b = 0; b < B; ++b) {
    self[b] = ~0ull;
    if (!new_ids) continue;
    const Id128 key = load_id(new_ids + 16 * b);
    const uint32_t* sh_of = S->id2shard.find(key);
    if (!sh_of) continue;
    cx_index* c = S->sh[*sh_of];
    if (const uint32_t* row = c->id2row.find(key)) self[b] = c->h_seq[*row];
  }
  CU(cudaMemcpyAsync(d + o_self, self, B * 8, cudaMemcpyHostToDevice, mw->stream));
  shard_autolink_kernel<<<(unsigned)((B + 127) / 128), 128, 0, mw->stream>>>(
      f.m.seq, f.m.score, f.m.ids, f.m.n, (const uint64_t*)(d + o_self), (uint32_t)B, (uint32_t)kk, threshold, me,
      (float*)(d + o_os), (uint8_t*)(d + o_oi), (uint32_t*)(d + o_on));
  CU(cudaGetLastError());
  S->sh[0]->launches += 1;
  CU(cudaMemcpyAsync(hp + o_os, d + o_os, total - o_os, cudaMemcpyDeviceToHost, mw->stream));
  CU(cudaStreamSynchronize(mw->stream));
  memcpy(out_score, hp + o_os, B * me * 4);
  memcpy(out_to_ids, hp + o_oi, B * me * 16);
  memcpy(out_n, hp + o_on, B * 4);
  S->sh[0]->d2h += B * me * 20 + B * 4;
  return CX_OK;
}

// DedupScanner::scan (linker/dedup.rs:65-127) over the sharded corpus: blocks of one shard's rows are the
// queries, every shard is scanned for partners inserted AFTER the query (each unordered pair once, the
// node itself never), per-node lists are merged on devices[0] and compacted there.  The blocks of the
// different shards are finally interleaved by insertion order of the first node, which is the order a
// single index reports.
cx_status shard_dedup_scan(cx_index* h, float threshold, uint32_t per_node_cap, uint64_t max_pairs, uint8_t* out_a_ids,
                           uint8_t* out_b_ids, float* out_score, uint64_t* out_n, uint64_t* out_total) {
  ShardSet* S = h->shards;
  const uint32_t W = S->W();
  if (shard_len(h) < 2 || !per_node_cap || threshold != threshold) return CX_OK;
  for (cx_index* c : S->sh) {
    cx_status e = settle(c);
    if (e != CX_OK) return e;
  }
  uint64_t n_rows = 0;
  for (cx_index* c : S->sh) n_rows += c->n_rows;
  const uint32_t kd = (uint32_t)std::min<uint64_t>(per_node_cap, n_rows);
  const uint64_t QB = (uint64_t)S->sh[0]->sm_count * 128u;
  struct Rec {
    uint64_t seq_a;
    uint32_t shard, row;
    float score;
    uint8_t b_id[16];
  };
  std::vector<std::vector<Rec>> stream(W);  // per query shard, already in insertion order of a
  uint64_t n_tot = 0;
  for (uint32_t qs = 0; qs < W; ++qs) {
    cx_index* qc = S->sh[qs];
    for (uint64_t r0 = 0; r0 < qc->n_rows; r0 += QB) {
      const uint64_t B = std::min<uint64_t>(QB, qc->n_rows - r0);
      FanOut f(S);
      f.B = B;
      f.k = kd;
      cx_status st = f.ml.init();
      if (st != CX_OK) return st;
      f.m = carve_merged(nullptr, W, B, kd, kd);
      // + offsets [B+1], dead flags [B], totals [B], records [B*kd][6]
      const size_t o_off = align_up(f.m.dev_bytes, 256), o_dead = align_up(o_off + (B + 1) * 4, 256),
                   o_rec = align_up(o_dead + B * 4, 256);
      const uint64_t rec_cap = std::min<uint64_t>(B * kd, max_pairs ? max_pairs : 1);
      const size_t dev_total = o_rec + rec_cap * 24;
      const size_t h_rec = 0, h_dead = align_up(rec_cap * 24, 256), h_cnt = align_up(h_dead + B * 4, 256);
      st = f.ml.ensure(dev_total, h_cnt + 256);
      if (st != CX_OK) return st;
      f.m = carve_merged(f.ml.ws->d, W, B, kd, kd);
      MergeWs* mw = f.ml.ws;
      char* dm = (char*)mw->d;
      char* hm = (char*)mw->hp;
      std::vector<std::vector<uint64_t>> totals(W, std::vector<uint64_t>(B, 0));
      const uint64_t* q_seq_h = qc->h_seq.data() + r0;
      st = for_shards(S, [&](uint32_t s) -> cx_status {
        cx_index* c = S->sh[s];
        ShardCall& sc = f.calls[s];
        sc.kd = kd;
        cx_status e = shard_call_init(c, &sc, B, c->ld, kd, !S->peer[s], true);
        if (e != CX_OK) return e;
        cudaStream_t stx = sc.outer->ws->stream;
        if (c->n_live == 0) {
          CU(cudaMemsetAsync(sc.n, 0, B * 4, stx));
          return shard_pack(S, s, sc, mw, f.m.gathered, B, kd, nullptr);
        }
        // the query block: rows of shard qs, padded rows as they sit in its store
        const float* dq = qc->dE + (size_t)r0 * qc->ld;
        if (c->device != qc->device) {
          CU(cudaMemcpyPeerAsync(sc.dQ, c->device, dq, qc->device, B * (size_t)qc->ld * 4, stx));
          dq = sc.dQ;
        } else if (c != qc) {
          CU(cudaMemcpyAsync(sc.dQ, dq, B * (size_t)qc->ld * 4, cudaMemcpyDeviceToDevice, stx));
          dq = sc.dQ;
        }
        CU(cudaMemcpyAsync(sc.q_seq, q_seq_h, B * 8, cudaMemcpyHostToDevice, stx));
        CU(cudaStreamSynchronize(stx));
        PairRules pr;
        pr.d_q_seq = sc.q_seq;
        pr.h_q_seq = q_seq_h;
        // rows inserted before the block's first node cannot be partners of any of its nodes
        pr.first_row = (uint32_t)(std::upper_bound(c->h_seq.begin(), c->h_seq.begin() + c->n_rows, q_seq_h[0]) -
                                  c->h_seq.begin());
        if (pr.first_row >= c->n_rows) {
          CU(cudaMemsetAsync(sc.n, 0, B * 4, stx));
          return shard_pack(S, s, sc, mw, f.m.gathered, B, kd, nullptr);
        }
        sc.kd = (uint32_t)std::min<uint64_t>(kd, c->n_rows);
        e = index_threshold_device(c, dq, c->dim, c->ld, B, threshold, nullptr, sc.kd, sc.rows, sc.score, sc.dist,
                                   nullptr, sc.n, nullptr, totals[s].data(), &pr);
        if (e != CX_OK) return e;
        return shard_pack(S, s, sc, mw, f.m.gathered, B, kd, nullptr);
      });
      if (st != CX_OK) return st;
      st = launch_merge(S, mw, f.m, B, kd, kd, nullptr, f.m.score, f.m.dist, f.m.ids, nullptr, f.m.n);
      if (st != CX_OK) return st;
      // removed nodes search for nothing (dedup.rs:73-76)
      uint32_t* dead = (uint32_t*)(hm + h_dead);
      for (uint64_t i = 0; i < B; ++i) dead[i] = qc->h_meta[r0 + i] & META_DEAD;
      CU(cudaSetDevice(S->merge_dev));
      CU(cudaMemcpyAsync(dm + o_dead, dead, B * 4, cudaMemcpyHostToDevice, mw->stream));
      launch_compact_offsets(f.m.n, (const uint32_t*)(dm + o_dead), (uint32_t)B, (uint32_t*)(dm + o_off), mw->stream);
      shard_pair_write_kernel<<<(unsigned)((B + 7) / 8), 256, 0, mw->stream>>>(
          f.m.ids, f.m.score, f.m.n, (const uint32_t*)(dm + o_off), (uint32_t)B, kd, (uint32_t*)(dm + o_rec),
          (uint32_t)rec_cap);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(hm + h_cnt, dm + o_off + B * 4, 4, cudaMemcpyDeviceToHost, mw->stream));
      CU(cudaStreamSynchronize(mw->stream));
      const uint64_t got = std::min<uint64_t>(*(uint32_t*)(hm + h_cnt), rec_cap);
      if (got) {
        CU(cudaMemcpyAsync(hm + h_rec, dm + o_rec, got * 24, cudaMemcpyDeviceToHost, mw->stream));
        CU(cudaStreamSynchronize(mw->stream));
        S->sh[0]->d2h += got * 24;
        const uint32_t* rec = (const uint32_t*)(hm + h_rec);
        for (uint64_t j = 0; j < got; ++j) {
          Rec r;
          r.row = (uint32_t)(r0 + rec[6 * j]);
          r.shard = qs;
          r.seq_a = qc->h_seq[r.row];
          memcpy(&r.score, &rec[6 * j + 1], 4);
          memcpy(r.b_id, &rec[6 * j + 2], 16);
          stream[qs].push_back(r);
        }
      }
      for (uint64_t i = 0; i < B; ++i) {
        if (dead[i]) continue;
        for (uint32_t s = 0; s < W; ++s) n_tot += totals[s][i];
      }
    }
  }
  // interleave the shards' streams by insertion order of a (each stream is already ordered)
  std::vector<size_t> at(W, 0);
  uint64_t n_out = 0;
  while (n_out < max_pairs) {
    uint32_t best = W;
    for (uint32_t s = 0; s < W; ++s)
      if (at[s] < stream[s].size() && (best == W || stream[s][at[s]].seq_a < stream[best][at[best]].seq_a)) best = s;
    if (best == W) break;
    const Rec& r = stream[best][at[best]++];
    memcpy(out_a_ids + 16 * n_out, S->sh[r.shard]->h_ids.data() + 16 * (size_t)r.row, 16);
    memcpy(out_b_ids + 16 * n_out, r.b_id, 16);
    out_score[n_out] = r.score;
    ++n_out;
  }
  *out_n = n_out;
  if (out_total) *out_total = n_tot;
  return CX_OK;
}

}  // namespace cx

using namespace cx;

extern "C" cx_status cx_index_create_sharded(uint32_t dimension, const int* devices, uint32_t n_devices,
                                             cx_index** out) {
  cx::CallerDevice keep_callers_device;
  if (!out) return fail(CX_ERR_VALIDATION, "out is null");
  *out = nullptr;
  if (!devices || n_devices == 0 || n_devices > 64) return fail(CX_ERR_VALIDATION, "need 1..64 devices");
  std::unique_ptr<cx_index, void (*)(cx_index*)> h(new cx_index(), cx_index_destroy);
  h->dim = dimension;
  h->ld = (uint32_t)align_up(dimension, 4);
  h->ld16 = (uint32_t)align_up(dimension, 64);
  h->device = devices[0];
  ShardSet* S = new ShardSet();
  h->shards = S;
  S->merge_dev = devices[0];
  for (uint32_t s = 0; s < n_devices; ++s) {
    cx_index* c = nullptr;
    cx_status st = index_create(dimension, devices[s], /*with_seq=*/true, &c);
    if (st != CX_OK) return st;
    S->sh.push_back(c);
    S->dev.push_back(devices[s]);
  }
  // producers store into the merge device's memory where the hardware allows it
  S->peer.assign(n_devices, 0);
  for (uint32_t s = 0; s < n_devices; ++s) {
    if (devices[s] == S->merge_dev) {
      S->peer[s] = 1;
      continue;
    }
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, devices[s], S->merge_dev) != cudaSuccess) can = 0;
    if (can) {
      CU(cudaSetDevice(devices[s]));
      cudaError_t e = cudaDeviceEnablePeerAccess(S->merge_dev, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
      else if (e != cudaSuccess) {
        (void)cudaGetLastError();
        can = 0;
      }
    }
    S->peer[s] = (char)can;
  }
  CU(cudaSetDevice(S->merge_dev));
  for (uint32_t s = 0; s < n_devices; ++s) {
    S->workers.emplace_back(new Worker());
    Worker* w = S->workers.back().get();
    const int d = devices[s];
    w->th = std::thread([w, d] { w->run(d); });
  }
  *out = h.release();
  return CX_OK;
}

extern "C" uint32_t cx_shard_count(const cx_index* h) {
  if (!h) return 0;
  return h->shards ? h->shards->W() : 1;
}

extern "C" cx_status cx_load_sharded(const char* path, const int* devices, uint32_t n_devices, cx_index** out) {
  cx::CallerDevice keep_callers_device;
  if (!path || !out) return fail(CX_ERR_VALIDATION, "null argument");
  struct Ctx {
    const int* dev;
    uint32_t n;
  } ctx{devices, n_devices};
  auto make = [](uint32_t dim, void* c, cx_index** o) {
    Ctx* x = (Ctx*)c;
    return cx_index_create_sharded(dim, x->dev, x->n, o);
  };
  return index_load_into(path, make, &ctx, out);
}
