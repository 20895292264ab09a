// cx_index.h -- internal host-side state shared by cx_index.cu (store, mutation,
// persistence) and cx_search.cu (search entry points).
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
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

// Per-call scratch: one stream + device/pinned buffers, recycled through a pool so
// concurrent searches (read-guard holders in the reference) never share state.
struct Workspace {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_sync = nullptr, ev_block = nullptr;
  void* d = nullptr;
  size_t d_bytes = 0;
  void* hp = nullptr;  // pinned
  size_t h_bytes = 0;
  // candidate-list state that must be zero between passes (cnt u32 + gtau u64 per query)
  uint32_t* d_cnt = nullptr;
  uint64_t* d_gtau = nullptr;
  size_t state_q = 0;
  bool state_dirty = false;

  cudaError_t ensure(size_t db, size_t hb);
  cudaError_t ensure_state(size_t nq);
  cudaError_t wait(bool blocking);
  ~Workspace();
};

}  // namespace cx

struct cx_index {
  int device = 0;
  int sm_count = 148;
  uint32_t dim = 0, ld = 0, ld16 = 0;
  uint64_t n_rows = 0, n_live = 0, cap = 0;
  float* dE = nullptr;
  float* dNorm = nullptr;
  float* dRnorm = nullptr;
  uint32_t* dMeta = nullptr;
  uint32_t* dAgent = nullptr;
  uint8_t* dIds = nullptr;
  void* dE16 = nullptr;
  uint32_t* d_irr = nullptr;       // device counter + pinned host mirror: rows whose norm under/overflowed
  uint32_t* h_irr = nullptr;       //   fp32 (cx_exact.cu); any such row sends every search to the exact path
  bool want_shadow = true;
  std::vector<uint8_t> h_ids;
  std::vector<uint32_t> h_meta, h_agent;
  std::unordered_map<cx::Id128, uint32_t, cx::Id128Hash> id2row;
  std::unordered_map<cx::Id128, std::pair<uint32_t, uint32_t>, cx::Id128Hash> orphan_meta;
  cx::Interner kinds, agents;
  cudaStream_t mut_stream = nullptr;
  std::mutex ws_mu;
  std::vector<cx::Workspace*> ws_free;
  // options
  int force_path = 0;
  uint32_t tensor_min_batch = 5;   // query groups at least this large go to the tensor pass
  uint32_t tensor_phase_growth = 0xFFFFFFFFu; // tensor pass: each scan phase covers this many times the rows seen before
                                              // (0/1 = one phase; 0xFFFFFFFF = auto: one phase for B <= 256 and k <= 16, else 8 for k <= 16, 4 above)
  int profile = 0;
  bool blocking_sync = false;      // search calls sleep on an event instead of spinning while the GPU works
  // stats
  std::atomic<uint64_t> launches{0}, q_stream{0}, q_tensor{0}, q_exact{0}, fallbacks{0}, h2d{0}, d2h{0};
  std::atomic<uint64_t> pass_ns{0}, pass_launches{0};

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

}  // namespace cx
