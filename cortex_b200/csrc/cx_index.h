// cx_index.h -- internal host-side state shared by cx_index.cu (store, mutation,
// persistence) and cx_search.cu (search entry points).
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cortex_gpu.h"
#include "cx_kernels.h"

namespace cx {

cx_status fail(cx_status st, const char* fmt, ...);
const char* last_error();

#define CU(expr)                                                                                     \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return cx::fail(CX_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__,  \
                      __LINE__, #expr);                                                              \
  } while (0)

// The library switches the calling thread's current device as it works (cudaSetDevice in the entry points, the
// shards of a multi-device handle one after the other).  Every C-ABI entry point puts the caller's device back
// on return, so the library can share a thread with other CUDA users (a Rust host with its own kernels, torch).
struct CallerDevice {
  int dev = -1;
  CallerDevice() {
    if (cudaGetDevice(&dev) != cudaSuccess) {
      dev = -1;
      (void)cudaGetLastError();
    }
  }
  ~CallerDevice() {
    if (dev >= 0) (void)cudaSetDevice(dev);
  }
  CallerDevice(const CallerDevice&) = delete;
  CallerDevice& operator=(const CallerDevice&) = delete;
};

struct Id128 {
  uint64_t a, b;
  bool operator==(const Id128& o) const { return a == o.a && b == o.b; }
};
struct Id128Hash {
  size_t operator()(const Id128& k) const {
    uint64_t h = k.a * 0x9E3779B97F4A7C15ull ^ (k.b + 0xC2B2AE3D27D4EB4Full);
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    return (size_t)(h ^ (h >> 32));
  }
};
inline Id128 load_id(const uint8_t* p) {
  Id128 k;
  memcpy(&k.a, p, 8);
  memcpy(&k.b, p + 8, 8);
  return k;
}

// id -> row.  256 independent tables picked by the hash, so that a table never rehashes more than 1/256th of
// the ids at once (one big table stalls an insert for ~0.2 s when it rehashes at a few million rows).
class IdMap {
 public:
  uint32_t* find(const Id128& k) {
    auto& m = sub_[slot(k)];
    auto it = m.find(k);
    return it == m.end() ? nullptr : &it->second;
  }
  const uint32_t* find(const Id128& k) const {
    const auto& m = sub_[slot(k)];
    auto it = m.find(k);
    return it == m.end() ? nullptr : &it->second;
  }
  bool count(const Id128& k) const { return find(k) != nullptr; }
  bool emplace(const Id128& k, uint32_t row) { return sub_[slot(k)].emplace(k, row).second; }
  void erase(const Id128& k) { sub_[slot(k)].erase(k); }
  void clear() {
    for (auto& m : sub_) m.clear();
  }
  size_t size() const {
    size_t n = 0;
    for (const auto& m : sub_) n += m.size();
    return n;
  }

 private:
  static size_t slot(const Id128& k) { return (Id128Hash()(k) >> 24) & 255u; }
  std::unordered_map<Id128, uint32_t, Id128Hash> sub_[256];
};

struct Interner {
  std::vector<std::string> strs;
  std::unordered_map<std::string, uint32_t> map;
  uint32_t intern(const std::string& s) {
    auto it = map.find(s);
    if (it != map.end()) return it->second;
    uint32_t id = (uint32_t)strs.size();
    strs.push_back(s);
    map.emplace(s, id);
    return id;
  }
  bool find(const std::string& s, uint32_t* id) const {
    auto it = map.find(s);
    if (it == map.end()) return false;
    *id = it->second;
    return true;
  }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// A device array that grows in place.  With the driver's virtual-memory API (cuMemAddressReserve /
// cuMemCreate / cuMemMap) a range of addresses large enough for the biggest shard the device can hold
// is reserved once and physical chunks are mapped behind it as rows arrive: the base pointer never
// changes, nothing is copied and growth needs no second copy of the store.  Where that API is not
// available the array falls back to allocate-copy-free (the pointer then changes on growth).
struct VmArray {
  char* p = nullptr;
  size_t reserved = 0;   // bytes of address space (vmm only)
  size_t mapped = 0;     // bytes usable
  bool vmm = false;
  int device = 0;
  struct Chunk { unsigned long long handle; size_t off, size; };
  std::vector<Chunk> chunks;
  cx_status init(int device, size_t reserve_bytes);
  // make at least `bytes` usable; the fallback path preserves the first `used` bytes (copied on s, synchronised)
  cx_status ensure(size_t bytes, size_t used, cudaStream_t s);
  void shrink(size_t bytes);  // vmm only: give back whole chunks that lie entirely above `bytes`
  void destroy();
};
bool vmm_available();

// Per-call scratch: one stream + device/pinned buffers, recycled through a pool so
// concurrent searches (read-guard holders in the reference) never share state.
struct Workspace {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_sync = nullptr, ev_block = nullptr;
  void* d = nullptr;
  size_t d_bytes = 0;
  void* hp = nullptr;  // pinned
  size_t h_bytes = 0;
  // exact-path buffers (two key arrays of n_rows + radix-sort scratch): allocated the first time a
  // query of this workspace actually takes the exact path, not per search
  void* dx = nullptr;
  size_t dx_bytes = 0;
  // second device block for callers that stage their own inputs / outputs around a device search
  void* aux = nullptr;
  size_t aux_bytes = 0;
  void* aux_h = nullptr;  // pinned
  size_t aux_h_bytes = 0;
  // replayable launch sequence of the last device-resident search shape (cx_search.cu)
  cudaGraphExec_t graph = nullptr;
  uint32_t graph_nlaunch = 0;   // kernels one replay launches
  bool graph_broken = false;    // recording failed once on this workspace: plain launches from then on
  uint64_t graph_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t last_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // candidate-list state that must be zero between passes (cnt u32 + gtau u64 per query)
  uint32_t* d_cnt = nullptr;
  uint64_t* d_gtau = nullptr;
  size_t state_q = 0;
  bool state_dirty = false;

  cudaError_t ensure(size_t db, size_t hb);
  cudaError_t ensure_exact(size_t bytes);
  cudaError_t ensure_aux(size_t db, size_t hb);
  cudaError_t ensure_state(size_t nq);
  void drop_graph();
  cudaError_t wait(bool blocking);
  ~Workspace();
};

}  // namespace cx

namespace cx { struct ShardSet; }

struct cx_index {
  int device = 0;
  int sm_count = 148;
  uint32_t dim = 0, ld = 0, ld16 = 0;
  uint64_t n_rows = 0, n_live = 0;
  std::atomic<uint64_t> cap{0};    // rows the mapped store can hold
  // in-place growth runs AHEAD of need on a helper thread (mapping a gigabyte takes the driver a few hundred
  // milliseconds; no insert should ever wait for that)
  std::thread grow_thread;
  cx_status grow_status = CX_OK;
  std::string grow_error;
  std::atomic<uint64_t> grow_ns{0}, grow_ns_max{0}, grow_waits{0};
  // the store: growable device arrays and the raw pointers into them
  cx::VmArray aE, aNorm, aRnorm, aMeta, aAgent, aIds, aE16, aSeq;
  float* dE = nullptr;
  float* dNorm = nullptr;
  float* dRnorm = nullptr;
  uint32_t* dMeta = nullptr;
  uint32_t* dAgent = nullptr;
  uint8_t* dIds = nullptr;
  void* dE16 = nullptr;
  uint64_t* dSeq = nullptr;        // shards of a multi-device index only: global insertion number of every row
  bool with_seq = false;
  uint64_t max_rows = 0;           // rows the reserved address ranges can hold
  size_t total_mem = 0;
  bool store_ready = false;        // address ranges reserved (first insert / reserve)
  void* stage_h = nullptr;         // pinned staging for the side arrays of an insert
  size_t stage_bytes = 0;
  uint32_t* d_irr = nullptr;       // device counter + pinned host mirror: LIVE rows whose norm under/overflowed
  uint32_t* h_irr = nullptr;       //   fp32 (cx_exact.cu); any such row sends every search to the exact path
  bool want_shadow = true;
  std::vector<uint8_t> h_ids;
  std::vector<uint32_t> h_meta, h_agent;
  std::vector<uint64_t> h_seq;
  cx::IdMap id2row;
  std::unordered_map<cx::Id128, std::pair<uint32_t, uint32_t>, cx::Id128Hash> orphan_meta;
  cx::Interner kinds, agents;
  cudaStream_t mut_stream = nullptr;
  cudaEvent_t ev_mut = nullptr;          // recorded on mut_stream after every mutation
  std::atomic<bool> mut_dirty{false};    // a mutation was enqueued that no search has waited for yet
  std::mutex ws_mu;
  std::vector<cx::Workspace*> ws_free;
  // options
  int force_path = 0;
  bool stream_bf16 = true;         // batches of up to four queries stream the bf16 shadow (K1, HALF variant)
  uint32_t tensor_min_batch = 3;   // query groups at least this large go to the tensor pass (measured: K1 over the
                                   // bf16 shadow wins at B = 1 and 2, the tensor pass from B = 3)
  uint32_t tensor_phase_growth = 0xFFFFFFFFu; // tensor pass: each scan phase covers this many times the rows seen before
                                              // (0/1 = one phase; 0xFFFFFFFF = auto: one phase for B <= 256 and k <= 16, else 8 for k <= 16, 6 above)
  uint32_t tensor_sample_tiles = 0;  // row tiles sampled for the cut-off bootstrap (0 = auto)
  int profile = 0;
  bool blocking_sync = false;      // search calls sleep on an event instead of spinning while the GPU works
  bool use_graphs = true;          // replay repeated device-resident search shapes as one CUDA graph launch
  cx::TensorTuning tensor_tune;    // CTA form / epilogue warps of the tensor pass (per index, read-only while searching)
  // stats
  std::atomic<uint64_t> launches{0}, q_stream{0}, q_stream16{0}, q_tensor{0}, q_exact{0}, fallbacks{0}, h2d{0}, d2h{0};
  std::atomic<uint64_t> pass_ns{0}, pass_launches{0}, graph_launches{0}, grow_events{0};

  // non-null: this handle is a row-sharded index over several devices (cx_sharded.cu); none of the
  // single-device fields above is used and every entry point forwards to the shard set
  cx::ShardSet* shards = nullptr;

  uint32_t n_irregular() const { return h_irr ? *h_irr : 0u; }

  cx::StoreView view() const {
    cx::StoreView v;
    v.E = dE;
    v.norm = dNorm;
    v.rnorm = dRnorm;
    v.meta = dMeta;
    v.agent = dAgent;
    v.ids = dIds;
    v.E16 = dE16;
    v.n_rows = (uint32_t)n_rows;
    v.dim = dim;
    v.ld = ld;
    v.ld16 = ld16;
    return v;
  }
};

namespace cx {

struct WsLease {
  cx_index* h;
  Workspace* ws;
  explicit WsLease(cx_index* h_) : h(h_), ws(nullptr) {
    std::lock_guard<std::mutex> g(h->ws_mu);
    if (!h->ws_free.empty()) {
      ws = h->ws_free.back();
      h->ws_free.pop_back();
    }
  }
  cudaError_t init();
  ~WsLease() {
    if (!ws) return;
    std::lock_guard<std::mutex> g(h->ws_mu);
    h->ws_free.push_back(ws);
  }
};

enum { PATH_AUTO = 0, PATH_STREAM = 1, PATH_TENSOR = 2, PATH_EXACT = 3 };

// wait (on the host) for mutations that were enqueued but not yet waited for: after it the store, the
// irregular-row count and every host mirror are current
cx_status settle(cx_index* h);

// The dedup scanner's pair rules (linker/dedup.rs:91-105) for a threshold scan whose queries are rows of an index
struct PairRules {
  const uint32_t* d_self_rows = nullptr;  // single index: device [B], the row each query is; it is skipped
  uint32_t self_base = 0;                 // ... and equals self_base + b for query b
  bool upper_only = false;                // ... and only rows above it are kept (each unordered pair once)
  uint32_t first_row = 0;                 // rows below this one cannot qualify (the scan starts at its tile)
  const uint64_t* d_q_seq = nullptr;      // multi-device: global insertion number of each query (device [B]); a row
  const uint64_t* h_q_seq = nullptr;      //   qualifies when its own number is larger.  Host copy of the same.
};

// single-device internals shared with the multi-device index (cx_sharded.cu)
cx_status index_threshold_device(cx_index* h, const float* d_queries, uint32_t qlen, uint32_t q_stride, uint64_t B,
                                 float threshold, const cx_filter* filter, uint64_t cap, uint32_t* d_rows,
                                 float* d_score, float* d_dist, uint8_t* d_ids, uint32_t* d_n, uint32_t* d_total,
                                 uint64_t* h_total, const PairRules* pair);
// cx_search_batch_device(_begin) for queries of any length ([B][qlen], the reference never checks it)
cx_status index_search_device(cx_index* h, const float* d_queries, uint32_t qlen, uint64_t B, uint64_t k,
                              const cx_filter* filter, uint32_t* d_out_rows, float* d_out_score, float* d_out_distance,
                              uint8_t* d_out_ids, uint32_t* d_out_n, void* stream, void** ticket);
cx_status index_create(uint32_t dimension, int device, bool with_seq, cx_index** out);
cx_status index_prepare_store(cx_index* h);
cx_status index_insert(cx_index* h, const uint8_t* ids, const float* rows, uint64_t n, uint32_t len, bool on_device,
                       const uint64_t* seq);
cx_status index_remove(cx_index* h, const uint8_t id[16]);
cx_status index_set_metadata(cx_index* h, const uint8_t id[16], const char* kind, const char* agent);
cx_status index_rebuild(cx_index* h);
cx_status index_set_option(cx_index* h, const char* key, int64_t value);
cx_status index_save_vectors(cx_index* h, FILE* fp, uint64_t r0, uint64_t r1);
uint64_t index_meta_count(const cx_index* h);
bool index_save_meta(const cx_index* h, FILE* fp);
cx_status index_load_into(const char* path, cx_status (*make)(uint32_t, void*, cx_index**), void* ctx, cx_index** out);
void index_add_stats(const cx_index* h, cx_stats* out);

// multi-device index (cx_sharded.cu): every entry point of the C ABI forwards here when h->shards is set
void shard_destroy(cx_index* h);
cx_status shard_reserve(cx_index* h, uint64_t n_rows);
cx_status shard_insert(cx_index* h, const uint8_t* ids, const float* rows, uint64_t n, uint32_t len, bool on_device);
cx_status shard_remove(cx_index* h, const uint8_t id[16]);
cx_status shard_set_metadata(cx_index* h, const uint8_t id[16], const char* kind, const char* agent);
uint64_t shard_len(const cx_index* h);
cx_status shard_rebuild(cx_index* h);
cx_status shard_save(cx_index* h, const char* path);
cx_status shard_stats(cx_index* h, cx_stats* out);
cx_status shard_set_option(cx_index* h, const char* key, int64_t value);
cx_status shard_search_host(cx_index* h, const float* queries, uint64_t B, uint32_t qlen, uint64_t k,
                            const cx_filter* filter, bool threshold_mode, float threshold, uint8_t* out_ids,
                            float* out_score, float* out_dist, uint64_t* out_n, uint64_t* out_total);
cx_status shard_dedup_scan(cx_index* h, float threshold, uint32_t per_node_cap, uint64_t max_pairs, uint8_t* out_a_ids,
                           uint8_t* out_b_ids, float* out_score, uint64_t* out_n, uint64_t* out_total);
cx_status shard_autolink_batch(cx_index* h, const uint8_t* new_ids, const float* embeddings, uint64_t B, uint64_t k,
                               float threshold, uint32_t max_edges_per_node, uint8_t* out_to_ids, float* out_score,
                               uint32_t* out_n);
cx_status shard_search_device(cx_index* h, const float* d_queries, uint64_t B, uint64_t k, const cx_filter* filter,
                              uint32_t* d_out_rows, float* d_out_score, float* d_out_distance, uint8_t* d_out_ids,
                              uint32_t* d_out_n, void* stream, void** ticket);
cx_status shard_search_device_end(cx_index* h, void* ticket, uint64_t* n_redone);

}  // namespace cx
