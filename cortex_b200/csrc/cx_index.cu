// cx_index.cu -- host side of the C ABI (include/cortex_gpu.h): the embedding
// store, per-call workspaces, path selection and the exported entry points.
//
// Mirrors HnswIndex (vector/index.rs:182-473) in exact-scan mode:
//   vectors  : HashMap<NodeId, Vec<f32>>  -> device matrix + id->row hash on the host
//   metadata : HashMap<NodeId, NodeMetadata> -> per-row meta/agent words + string tables
// There is no CPU compute path in this file: every score comes from a kernel.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cortex_gpu.h"
#include "cx_kernels.h"

using namespace cx;

// ------------------------------------------------------------------------------
static thread_local std::string g_err;

static cx_status fail(cx_status st, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return st;
}

#define CU(expr)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return fail(CX_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, \
                  __LINE__, #expr);                                                       \
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
static Id128 load_id(const uint8_t* p) {
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

// Per-call scratch: one stream + device/pinned buffers, recycled through a pool so
// concurrent searches (read-guard holders in the reference) never share state.
struct Workspace {
  cudaStream_t stream = nullptr;
  void* d = nullptr;
  size_t d_bytes = 0;
  void* hp = nullptr;  // pinned
  size_t h_bytes = 0;
  cudaError_t ensure(size_t db, size_t hb) {
    if (db > d_bytes) {
      if (d) cudaFree(d);
      d = nullptr;
      d_bytes = 0;
      size_t want = db + db / 4;
      cudaError_t e = cudaMalloc(&d, want);
      if (e != cudaSuccess) return e;
      d_bytes = want;
    }
    if (hb > h_bytes) {
      if (hp) cudaFreeHost(hp);
      hp = nullptr;
      h_bytes = 0;
      size_t want = hb + hb / 4;
      cudaError_t e = cudaMallocHost(&hp, want);
      if (e != cudaSuccess) return e;
      h_bytes = want;
    }
    return cudaSuccess;
  }
  ~Workspace() {
    if (d) cudaFree(d);
    if (hp) cudaFreeHost(hp);
    if (stream) cudaStreamDestroy(stream);
  }
};

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
  bool want_shadow = false;
  std::vector<uint8_t> h_ids;
  std::vector<uint32_t> h_meta, h_agent;
  std::unordered_map<Id128, uint32_t, Id128Hash> id2row;
  std::unordered_map<Id128, std::pair<uint32_t, uint32_t>, Id128Hash> orphan_meta;
  Interner kinds, agents;
  cudaStream_t mut_stream = nullptr;
  std::mutex ws_mu;
  std::vector<Workspace*> ws_free;
  // options
  int force_path = 0;
  uint32_t stream_max_batch = 1u << 30;
  // stats
  std::atomic<uint64_t> launches{0}, q_stream{0}, q_tensor{0}, q_exact{0}, fallbacks{0}, h2d{0}, d2h{0};
  std::atomic<uint64_t> pass_ns{0}, pass_launches{0};
  int profile = 0;

  StoreView view() const {
    StoreView v;
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

struct WsLease {
  cx_index* h;
  Workspace* ws;
  WsLease(cx_index* h_) : h(h_), ws(nullptr) {
    std::lock_guard<std::mutex> g(h->ws_mu);
    if (!h->ws_free.empty()) {
      ws = h->ws_free.back();
      h->ws_free.pop_back();
    }
  }
  cudaError_t init() {
    if (ws) return cudaSuccess;
    ws = new Workspace();
    return cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking);
  }
  ~WsLease() {
    if (!ws) return;
    std::lock_guard<std::mutex> g(h->ws_mu);
    h->ws_free.push_back(ws);
  }
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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

// ------------------------------------------------------------------------------
static cx_status grow(cx_index* h, uint64_t need) {
  if (need <= h->cap) return CX_OK;
  if (need > 0xFFFFFF00ull) return fail(CX_ERR_VALIDATION, "index shard limited to 2^32 rows");
  uint64_t ncap = h->cap ? h->cap * 2 : 1024;
  if (ncap < need) ncap = need;
  float *E = nullptr, *nm = nullptr, *rn = nullptr;
  uint32_t *me = nullptr, *ag = nullptr;
  uint8_t* ids = nullptr;
  void* e16 = nullptr;
  CU(cudaMalloc(&E, ncap * h->ld * sizeof(float)));
  CU(cudaMalloc(&nm, (ncap + 32) * sizeof(float)));  // +32: the scan reads norms in 16 B units
  CU(cudaMalloc(&rn, (ncap + 32) * sizeof(float)));
  CU(cudaMalloc(&me, ncap * sizeof(uint32_t)));
  CU(cudaMalloc(&ag, ncap * sizeof(uint32_t)));
  CU(cudaMalloc(&ids, ncap * 16));
  if (h->want_shadow) CU(cudaMalloc(&e16, ncap * h->ld16 * 2));
  cudaStream_t s = h->mut_stream;
  if (h->n_rows) {
    CU(cudaMemcpyAsync(E, h->dE, h->n_rows * h->ld * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(nm, h->dNorm, h->n_rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(rn, h->dRnorm, h->n_rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(me, h->dMeta, h->n_rows * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(ag, h->dAgent, h->n_rows * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(ids, h->dIds, h->n_rows * 16, cudaMemcpyDeviceToDevice, s));
    if (e16 && h->dE16)
      CU(cudaMemcpyAsync(e16, h->dE16, h->n_rows * h->ld16 * 2, cudaMemcpyDeviceToDevice, s));
  }
  CU(cudaStreamSynchronize(s));
  cudaFree(h->dE);
  cudaFree(h->dNorm);
  cudaFree(h->dRnorm);
  cudaFree(h->dMeta);
  cudaFree(h->dAgent);
  cudaFree(h->dIds);
  if (h->dE16) cudaFree(h->dE16);
  h->dE = E;
  h->dNorm = nm;
  h->dRnorm = rn;
  h->dMeta = me;
  h->dAgent = ag;
  h->dIds = ids;
  h->dE16 = e16;
  h->cap = ncap;
  h->h_ids.reserve(ncap * 16);
  return CX_OK;
}

static uint32_t meta_word(bool has, uint32_t kind) {
  return has ? (META_HAS | (kind & META_KIND_MASK)) : 0u;
}

extern "C" cx_status cx_index_create(uint32_t dimension, int device, cx_index** out) {
  if (!out) return fail(CX_ERR_VALIDATION, "out is null");
  *out = nullptr;
  if (dimension == 0 || dimension > 65536) return fail(CX_ERR_VALIDATION, "dimension %u out of range", dimension);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(CX_ERR_CUDA, "no CUDA device available (%s); cortex_gpu has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return fail(CX_ERR_VALIDATION, "device %d out of range", device);
  CU(cudaSetDevice(device));
  std::unique_ptr<cx_index> h(new cx_index());
  h->device = device;
  h->dim = dimension;
  h->ld = (uint32_t)align_up(dimension, 4);
  h->ld16 = (uint32_t)align_up(dimension, 64);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(CX_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                prop.minor);
  h->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&h->mut_stream, cudaStreamNonBlocking));
  *out = h.release();
  return CX_OK;
}

extern "C" void cx_index_destroy(cx_index* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (Workspace* w : h->ws_free) delete w;
  cudaFree(h->dE);
  cudaFree(h->dNorm);
  cudaFree(h->dRnorm);
  cudaFree(h->dMeta);
  cudaFree(h->dAgent);
  cudaFree(h->dIds);
  if (h->dE16) cudaFree(h->dE16);
  if (h->mut_stream) cudaStreamDestroy(h->mut_stream);
  delete h;
}

extern "C" cx_status cx_reserve(cx_index* h, uint64_t n_rows) {
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  CU(cudaSetDevice(h->device));
  return grow(h, n_rows);
}

extern "C" cx_status cx_insert_batch(cx_index* h, const uint8_t* ids, const float* rows, uint64_t n,
                                     uint32_t len) {
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (len != h->dim)
    return fail(CX_ERR_VALIDATION, "Embedding dimension mismatch: expected %u, got %u", h->dim, len);
  if (!n) return CX_OK;
  if (!ids || !rows) return fail(CX_ERR_VALIDATION, "null ids/rows");
  CU(cudaSetDevice(h->device));
  cx_status st = grow(h, h->n_rows + n);
  if (st != CX_OK) return st;
  cudaStream_t s = h->mut_stream;
  const uint64_t first_new = h->n_rows;
  std::vector<uint32_t> overwritten;
  uint64_t i = 0;
  while (i < n) {
    Id128 key = load_id(ids + 16 * i);
    auto it = h->id2row.find(key);
    if (it != h->id2row.end()) {  // HashMap::insert overwrite (index.rs:307)
      uint32_t r = it->second;
      CU(cudaMemcpyAsync(h->dE + (size_t)r * h->ld, rows + (size_t)i * len, len * sizeof(float),
                         cudaMemcpyHostToDevice, s));
      if (r < first_new) overwritten.push_back(r);
      ++i;
      continue;
    }
    // run of new ids
    uint64_t j = i;
    uint64_t r0 = h->n_rows;
    while (j < n) {
      Id128 kj = load_id(ids + 16 * j);
      if (h->id2row.count(kj)) break;
      uint32_t r = (uint32_t)h->n_rows++;
      h->id2row.emplace(kj, r);
      h->h_ids.insert(h->h_ids.end(), ids + 16 * j, ids + 16 * j + 16);
      uint32_t mw = 0, ag = 0;
      auto om = h->orphan_meta.find(kj);
      if (om != h->orphan_meta.end()) {
        mw = meta_word(true, om->second.first);
        ag = om->second.second;
        h->orphan_meta.erase(om);
      }
      h->h_meta.push_back(mw);
      h->h_agent.push_back(ag);
      h->n_live++;
      ++j;
    }
    uint64_t cnt = j - i;
    CU(cudaMemcpy2DAsync(h->dE + (size_t)r0 * h->ld, h->ld * sizeof(float), rows + (size_t)i * len,
                         len * sizeof(float), len * sizeof(float), cnt, cudaMemcpyHostToDevice, s));
    i = j;
  }
  const uint64_t n_new = h->n_rows - first_new;
  h->h2d += n * len * sizeof(float);
  if (n_new) {
    CU(cudaMemcpyAsync(h->dIds + first_new * 16, h->h_ids.data() + first_new * 16, n_new * 16,
                       cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->dMeta + first_new, h->h_meta.data() + first_new, n_new * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->dAgent + first_new, h->h_agent.data() + first_new, n_new * 4, cudaMemcpyHostToDevice, s));
    launch_prepare_rows(h->dE, h->dNorm, h->dRnorm, h->dE16, h->dim, h->ld, h->ld16, (uint32_t)first_new,
                        (uint32_t)n_new, s);
    h->launches += h->dE16 ? 2 : 1;
  }
  for (uint32_t r : overwritten) {
    launch_prepare_rows(h->dE, h->dNorm, h->dRnorm, h->dE16, h->dim, h->ld, h->ld16, r, 1, s);
    h->launches += h->dE16 ? 2 : 1;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s));
  return CX_OK;
}

extern "C" cx_status cx_insert(cx_index* h, const uint8_t id[16], const float* embedding, uint32_t len) {
  return cx_insert_batch(h, id, embedding, 1, len);
}

extern "C" cx_status cx_remove(cx_index* h, const uint8_t id[16]) {
  if (!h || !id) return fail(CX_ERR_VALIDATION, "null argument");
  Id128 key = load_id(id);
  h->orphan_meta.erase(key);
  auto it = h->id2row.find(key);
  if (it == h->id2row.end()) return CX_OK;  // index.rs:316-323: never an error
  uint32_t r = it->second;
  h->id2row.erase(it);
  h->h_meta[r] |= META_DEAD;
  h->n_live--;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->dMeta + r, &h->h_meta[r], 4, cudaMemcpyHostToDevice, h->mut_stream));
  CU(cudaStreamSynchronize(h->mut_stream));
  return CX_OK;
}

extern "C" cx_status cx_set_metadata(cx_index* h, const uint8_t id[16], const char* kind, const char* agent) {
  if (!h || !id || !kind || !agent) return fail(CX_ERR_VALIDATION, "null argument");
  uint32_t k = h->kinds.intern(kind);
  if (k > META_KIND_MASK) return fail(CX_ERR_VALIDATION, "more than 256 distinct node kinds");
  uint32_t a = h->agents.intern(agent);
  Id128 key = load_id(id);
  auto it = h->id2row.find(key);
  if (it == h->id2row.end()) {  // metadata map is independent of the vector map (index.rs:219-222)
    h->orphan_meta[key] = {k, a};
    return CX_OK;
  }
  uint32_t r = it->second;
  h->h_meta[r] = meta_word(true, k) | (h->h_meta[r] & META_DEAD);
  h->h_agent[r] = a;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->dMeta + r, &h->h_meta[r], 4, cudaMemcpyHostToDevice, h->mut_stream));
  CU(cudaMemcpyAsync(h->dAgent + r, &h->h_agent[r], 4, cudaMemcpyHostToDevice, h->mut_stream));
  CU(cudaStreamSynchronize(h->mut_stream));
  return CX_OK;
}

extern "C" uint64_t cx_len(const cx_index* h) { return h ? h->n_live : 0; }
extern "C" uint32_t cx_dimension(const cx_index* h) { return h ? h->dim : 0; }

extern "C" cx_status cx_rebuild(cx_index* h) {
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (h->n_live == h->n_rows) return CX_OK;
  CU(cudaSetDevice(h->device));
  cudaStream_t s = h->mut_stream;
  std::vector<uint32_t> live;
  live.reserve(h->n_live);
  for (uint32_t r = 0; r < h->n_rows; ++r)
    if (!(h->h_meta[r] & META_DEAD)) live.push_back(r);
  const uint64_t nl = live.size();
  uint64_t ncap = nl > 1024 ? nl : 1024;
  float *E = nullptr, *nm = nullptr, *rn = nullptr;
  uint32_t *me = nullptr, *ag = nullptr, *dlive = nullptr;
  uint8_t* ids = nullptr;
  void* e16 = nullptr;
  CU(cudaMalloc(&E, ncap * h->ld * sizeof(float)));
  CU(cudaMalloc(&nm, (ncap + 32) * 4));
  CU(cudaMalloc(&rn, (ncap + 32) * 4));
  CU(cudaMalloc(&me, ncap * 4));
  CU(cudaMalloc(&ag, ncap * 4));
  CU(cudaMalloc(&ids, ncap * 16));
  if (h->dE16) CU(cudaMalloc(&e16, ncap * h->ld16 * 2));
  if (nl) {
    CU(cudaMalloc(&dlive, nl * 4));
    CU(cudaMemcpyAsync(dlive, live.data(), nl * 4, cudaMemcpyHostToDevice, s));
    launch_gather_rows(h->view(), E, nm, rn, me, ag, ids, e16, dlive, (uint32_t)nl, s);
    h->launches += 1;
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(s));
  if (dlive) cudaFree(dlive);
  cudaFree(h->dE);
  cudaFree(h->dNorm);
  cudaFree(h->dRnorm);
  cudaFree(h->dMeta);
  cudaFree(h->dAgent);
  cudaFree(h->dIds);
  if (h->dE16) cudaFree(h->dE16);
  h->dE = E;
  h->dNorm = nm;
  h->dRnorm = rn;
  h->dMeta = me;
  h->dAgent = ag;
  h->dIds = ids;
  h->dE16 = e16;
  h->cap = ncap;
  std::vector<uint8_t> nids(nl * 16);
  std::vector<uint32_t> nmeta(nl), nagent(nl);
  h->id2row.clear();
  for (uint64_t i = 0; i < nl; ++i) {
    memcpy(&nids[i * 16], &h->h_ids[(size_t)live[i] * 16], 16);
    nmeta[i] = h->h_meta[live[i]];
    nagent[i] = h->h_agent[live[i]];
    h->id2row.emplace(load_id(&nids[i * 16]), (uint32_t)i);
  }
  h->h_ids.swap(nids);
  h->h_meta.swap(nmeta);
  h->h_agent.swap(nagent);
  h->n_rows = nl;
  h->n_live = nl;
  return CX_OK;
}

// ------------------------------------------------------------------------------
// search machinery

struct FilterHost {
  DevFilter dev;
  std::vector<uint32_t> excl_rows;
  bool active = false;
};

static void build_filter(const cx_index* h, const cx_filter* f, FilterHost* out) {
  DevFilter& d = out->dev;
  memset(&d, 0, sizeof d);
  d.agent = AGENT_NONE;
  if (!f) return;
  out->active = true;
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
      auto it = h->id2row.find(load_id(f->exclude_ids + 16 * i));
      if (it != h->id2row.end()) out->excl_rows.push_back(it->second);
    }
  }
}

static uint32_t keep_count(uint32_t k, uint32_t G) {
  uint32_t margin = k / 4;
  if (margin < 16) margin = 16;
  if (margin > 32) margin = 32;
  uint32_t KP = k + margin;
  while (KP > k && (uint64_t)KP * G > 16384) --KP;
  return KP;
}

// eps bound for the fp32 streaming pass: |approx - reference| cosine
static float eps_stream(uint32_t dim) { return (2.1f * (float)dim + 16.0f) * 5.9604645e-8f; }

struct SearchBufs {
  float* dQ;
  float* qnorm;
  float* rqnorm;
  uint32_t* excl;
  uint64_t* cand_keys;
  uint64_t* cand_bound;
  uint64_t* gtau;
  uint32_t *rows, *n, *ok;
  float *score, *dist;
  uint8_t* ids;
  uint64_t *ekeys_a, *ekeys_b;
  void* sort_tmp;
  uint32_t* n_total;
  size_t sort_tmp_bytes;
};

enum { PATH_AUTO = 0, PATH_STREAM = 1, PATH_TENSOR = 2, PATH_EXACT = 3 };

// Core: queries already on device at dQ [B][ldq].  Results land in device result
// buffers (rows/score/dist/ids [B][kd], n[B]).  Returns after the stream is idle.
static cx_status run_search(cx_index* h, Workspace* ws, const FilterHost& fh, const SearchBufs& sb, uint64_t B,
                            uint32_t qlen, uint32_t ldq, uint32_t kd, bool threshold_mode, float threshold,
                            uint32_t* h_ok /* pinned, B */, uint64_t* h_total /* optional, threshold mode */) {
  cudaStream_t s = ws->stream;
  StoreView st = h->view();
  QueryView qv;
  qv.Q = sb.dQ;
  qv.qnorm = sb.qnorm;
  qv.rqnorm = sb.rqnorm;
  qv.nq = (uint32_t)B;
  qv.qlen = qlen;
  qv.ldq = ldq;
  DevFilter flt = fh.dev;
  flt.excl_rows = sb.excl;
  flt.n_excl = (uint32_t)fh.excl_rows.size();
  if (flt.n_excl)
    CU(cudaMemcpyAsync(sb.excl, fh.excl_rows.data(), flt.n_excl * 4, cudaMemcpyHostToDevice, s));
  launch_prepare_queries(sb.dQ, sb.qnorm, sb.rqnorm, (uint32_t)B, qlen, ldq, s);
  h->launches += 1;

  ResultView rv;
  rv.rows = sb.rows;
  rv.score = sb.score;
  rv.dist = sb.dist;
  rv.ids = sb.ids;
  rv.n = sb.n;
  rv.ok = sb.ok;
  rv.k = kd;

  const uint32_t G = stream_scan_groups(st.n_rows, h->sm_count);
  const uint32_t KP = keep_count(kd, G);
  bool fast = !threshold_mode && qlen == h->dim && st.n_rows >= 256 && kd <= 128 && KP >= kd &&
              stream_scan_smem(st.ld, 8, KP) != 0 && select_smem(G, KP, st.ld) <= 227 * 1024;
  if (h->force_path == PATH_EXACT) fast = false;
  if (h->force_path == PATH_STREAM && !fast && !threshold_mode)
    return fail(CX_ERR_VALIDATION, "force_path=stream but the call shape is not eligible");

  std::vector<uint32_t> redo;
  if (fast) {
    CandView cv;
    cv.keys = sb.cand_keys;
    cv.bound = sb.cand_bound;
    cv.gtau = sb.gtau;
    CU(cudaMemsetAsync(sb.gtau, 0, B * 8, s));
    cv.G = G;
    cv.KP = KP;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t n_pass = 0;
    if (h->profile) {
      CU(cudaEventCreate(&ev0));
      CU(cudaEventCreate(&ev1));
      CU(cudaEventRecord(ev0, s));
    }
    for (uint64_t q0 = 0; q0 < B; q0 += 8) {
      uint32_t nq = (uint32_t)(B - q0 < 8 ? B - q0 : 8);
      CU(launch_stream_scan(st, qv, (uint32_t)q0, nq, flt, cv, h->sm_count, s));
      h->launches += 1;
      ++n_pass;
    }
    if (h->profile) CU(cudaEventRecord(ev1, s));
    CU(launch_select_rescore(st, qv, 0, (uint32_t)B, cv, rv, eps_stream(h->dim), s));
    h->launches += 1;
    CU(cudaMemcpyAsync(h_ok, sb.ok, B * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (h->profile) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, ev0, ev1));
      h->pass_ns += (uint64_t)(ms * 1e6);
      h->pass_launches += n_pass;
      cudaEventDestroy(ev0);
      cudaEventDestroy(ev1);
    }
    for (uint64_t b = 0; b < B; ++b)
      if (!h_ok[b]) redo.push_back((uint32_t)b);
    h->q_stream += B - redo.size();
    h->fallbacks += redo.size();
  } else {
    redo.resize(B);
    for (uint64_t b = 0; b < B; ++b) redo[b] = (uint32_t)b;
  }

  // exact path (tiny indexes, huge k, threshold scans, unverifiable fast results)
  uint32_t min_ord = 1;  // every real key, NaN included (they sort last)
  if (threshold_mode) {
    if (threshold != threshold) min_ord = 0xFFFFFFFFu;  // score >= NaN is false
    else min_ord = ord_from_score(threshold > 0.0f ? threshold : 0.0f);
  }
  for (uint32_t b : redo) {
    launch_exact_keys(st, qv, b, flt, sb.ekeys_a, s);
    CU(exact_sort(sb.ekeys_a, sb.ekeys_b, st.n_rows, sb.sort_tmp, sb.sort_tmp_bytes, s));
    launch_exact_emit(st, qv, b, sb.ekeys_b, st.n_rows, kd, min_ord, sb.rows + (size_t)b * kd,
                      sb.score + (size_t)b * kd, sb.dist + (size_t)b * kd,
                      sb.ids ? sb.ids + (size_t)b * kd * 16 : nullptr, sb.n + b, sb.n_total, s);
    h->launches += 4;
    if (h_total) {
      uint32_t t32 = 0;
      CU(cudaMemcpyAsync(&t32, sb.n_total, 4, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      h_total[b] = t32;
    }
  }
  h->q_exact += redo.size();
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s));
  return CX_OK;
}

static size_t carve_bufs(void* base, const cx_index* h, uint64_t B, uint32_t ldq, uint32_t kd, uint32_t n_excl,
                         bool own_queries, bool own_results, SearchBufs* sb) {
  Carver c(base);
  const uint32_t n_rows = (uint32_t)h->n_rows;
  const uint32_t G = stream_scan_groups(n_rows, h->sm_count);
  const uint32_t KP = keep_count(kd, G);
  sb->dQ = own_queries ? c.take<float>(B * ldq) : nullptr;
  sb->qnorm = c.take<float>(B);
  sb->rqnorm = c.take<float>(B);
  sb->excl = c.take<uint32_t>(n_excl + 1);
  sb->cand_keys = c.take<uint64_t>((size_t)B * G * KP);
  sb->cand_bound = c.take<uint64_t>((size_t)B * G);
  sb->gtau = c.take<uint64_t>(B);
  sb->n = c.take<uint32_t>(B);
  sb->ok = c.take<uint32_t>(B);
  sb->n_total = c.take<uint32_t>(4);
  if (own_results) {
    sb->rows = c.take<uint32_t>(B * kd);
    sb->score = c.take<float>(B * kd);
    sb->dist = c.take<float>(B * kd);
    sb->ids = c.take<uint8_t>(B * kd * 16);
  }
  sb->ekeys_a = c.take<uint64_t>(n_rows);
  sb->ekeys_b = c.take<uint64_t>(n_rows);
  sb->sort_tmp_bytes = exact_sort_tmp_bytes(n_rows);
  sb->sort_tmp = c.take<char>(sb->sort_tmp_bytes);
  return align_up(c.off, 256);
}

static cx_status search_host(cx_index* h, const float* queries, uint64_t B, uint32_t qlen, uint64_t k,
                             const cx_filter* filter, bool threshold_mode, float threshold, uint8_t* out_ids,
                             float* out_score, float* out_dist, uint64_t* out_n, uint64_t* out_total) {
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (B && !queries) return fail(CX_ERR_VALIDATION, "null queries");
  if (!out_n) return fail(CX_ERR_VALIDATION, "null out_n");
  for (uint64_t b = 0; b < B; ++b) out_n[b] = 0;
  if (out_total)
    for (uint64_t b = 0; b < B; ++b) out_total[b] = 0;
  if (h->n_live == 0 || B == 0) return CX_OK;  // index.rs:331: empty -> Ok(vec![])
  if (qlen == 0) qlen = 0;
  CU(cudaSetDevice(h->device));
  uint64_t kd64 = k < h->n_rows ? k : h->n_rows;
  if (kd64 == 0) return CX_OK;
  const uint32_t kd = (uint32_t)kd64;
  const uint32_t ldq = (uint32_t)align_up(qlen > h->ld ? qlen : h->ld, 4);

  FilterHost fh;
  build_filter(h, filter, &fh);

  WsLease lease(h);
  CU(lease.init());
  Workspace* ws = lease.ws;
  SearchBufs sb;
  memset(&sb, 0, sizeof sb);
  size_t dbytes = carve_bufs(nullptr, h, B, ldq, kd, (uint32_t)fh.excl_rows.size(), true, true, &sb);
  // pinned: queries [B][ldq] + ok[B] + results
  size_t hq = align_up(B * ldq * 4, 256), hok = align_up(B * 4, 256);
  size_t hres = align_up(B * kd * 4, 256) * 2 + align_up(B * kd * 16, 256) + align_up(B * 4, 256);
  CU(ws->ensure(dbytes, hq + hok + hres));
  carve_bufs(ws->d, h, B, ldq, kd, (uint32_t)fh.excl_rows.size(), true, true, &sb);
  char* hp = (char*)ws->hp;
  float* hQ = (float*)hp;
  uint32_t* h_ok = (uint32_t*)(hp + hq);
  char* hr = hp + hq + hok;
  float* h_score = (float*)hr;
  float* h_dist = (float*)(hr + align_up(B * kd * 4, 256));
  uint8_t* h_ids = (uint8_t*)(hr + 2 * align_up(B * kd * 4, 256));
  uint32_t* h_n = (uint32_t*)(hr + 2 * align_up(B * kd * 4, 256) + align_up(B * kd * 16, 256));

  // stage queries zero padded to ldq
  for (uint64_t b = 0; b < B; ++b) {
    memcpy(hQ + b * ldq, queries + b * qlen, (size_t)qlen * 4);
    for (uint32_t d = qlen; d < ldq; ++d) hQ[b * ldq + d] = 0.0f;
  }
  cudaStream_t s = ws->stream;
  CU(cudaMemcpyAsync(sb.dQ, hQ, B * ldq * 4, cudaMemcpyHostToDevice, s));
  h->h2d += B * ldq * 4;

  std::vector<uint64_t> totals;
  if (threshold_mode) totals.assign(B, 0);
  cx_status stt = run_search(h, ws, fh, sb, B, qlen, ldq, kd, threshold_mode, threshold, h_ok,
                             threshold_mode ? totals.data() : nullptr);
  if (stt != CX_OK) return stt;

  CU(cudaMemcpyAsync(h_score, sb.score, B * kd * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_dist, sb.dist, B * kd * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_ids, sb.ids, B * kd * 16, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_n, sb.n, B * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  h->d2h += B * kd * 24 + B * 4;
  for (uint64_t b = 0; b < B; ++b) {
    uint32_t n = h_n[b];
    out_n[b] = n;
    if (out_total) out_total[b] = threshold_mode ? totals[b] : n;
    if (out_score) memcpy(out_score + b * k, h_score + b * kd, (size_t)n * 4);
    if (out_dist) memcpy(out_dist + b * k, h_dist + b * kd, (size_t)n * 4);
    if (out_ids) memcpy(out_ids + b * k * 16, h_ids + b * kd * 16, (size_t)n * 16);
  }
  return CX_OK;
}

extern "C" cx_status cx_search(cx_index* h, const float* query, uint32_t qlen, uint64_t k,
                               const cx_filter* filter, uint8_t* out_ids, float* out_score, float* out_distance,
                               uint64_t* out_n) {
  return search_host(h, query, 1, qlen, k, filter, false, 0.0f, out_ids, out_score, out_distance, out_n, nullptr);
}

extern "C" cx_status cx_search_batch(cx_index* h, const float* queries, uint64_t B, uint32_t qlen, uint64_t k,
                                     const cx_filter* filter, uint8_t* out_ids, float* out_score,
                                     float* out_distance, uint64_t* out_n) {
  return search_host(h, queries, B, qlen, k, filter, false, 0.0f, out_ids, out_score, out_distance, out_n,
                     nullptr);
}

extern "C" cx_status cx_search_threshold(cx_index* h, const float* query, uint32_t qlen, float threshold,
                                         const cx_filter* filter, uint64_t cap, uint8_t* out_ids,
                                         float* out_score, float* out_distance, uint64_t* out_n,
                                         uint64_t* out_total) {
  uint64_t total = 0;
  cx_status st = search_host(h, query, 1, qlen, cap, filter, true, threshold, out_ids, out_score, out_distance,
                             out_n, &total);
  if (out_total) *out_total = total;
  return st;
}

extern "C" cx_status cx_search_batch_device(cx_index* h, const float* d_queries, uint64_t B, uint64_t k,
                                            const cx_filter* filter, uint32_t* d_out_rows, float* d_out_score,
                                            float* d_out_distance, uint8_t* d_out_ids, uint32_t* d_out_n,
                                            void* stream) {
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (!d_queries || !d_out_rows || !d_out_score || !d_out_distance || !d_out_n)
    return fail(CX_ERR_VALIDATION, "null device buffer");
  if (B == 0) return CX_OK;
  CU(cudaSetDevice(h->device));
  cudaStream_t user = (cudaStream_t)stream;
  if (h->n_live == 0 || k == 0) {
    CU(cudaMemsetAsync(d_out_n, 0, B * 4, user));
    return CX_OK;
  }
  if (k > h->n_rows) return fail(CX_ERR_VALIDATION, "device search needs k <= rows in the shard");
  const uint32_t kd = (uint32_t)k;
  const uint32_t ldq = h->ld;
  FilterHost fh;
  build_filter(h, filter, &fh);
  WsLease lease(h);
  CU(lease.init());
  Workspace* ws = lease.ws;
  SearchBufs sb;
  memset(&sb, 0, sizeof sb);
  const bool own_q = h->dim != h->ld;
  size_t dbytes = carve_bufs(nullptr, h, B, ldq, kd, (uint32_t)fh.excl_rows.size(), own_q, false, &sb);
  CU(ws->ensure(dbytes, align_up(B * 4, 256)));
  carve_bufs(ws->d, h, B, ldq, kd, (uint32_t)fh.excl_rows.size(), own_q, false, &sb);
  // order after whatever produced the queries on the caller's stream
  cudaEvent_t ev;
  CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CU(cudaEventRecord(ev, user));
  CU(cudaStreamWaitEvent(ws->stream, ev, 0));
  if (own_q) {
    CU(cudaMemsetAsync(sb.dQ, 0, B * ldq * 4, ws->stream));
    CU(cudaMemcpy2DAsync(sb.dQ, ldq * 4, d_queries, h->dim * 4, h->dim * 4, B, cudaMemcpyDeviceToDevice,
                         ws->stream));
  } else {
    sb.dQ = const_cast<float*>(d_queries);
  }
  sb.rows = d_out_rows;
  sb.score = d_out_score;
  sb.dist = d_out_distance;
  sb.ids = d_out_ids;
  sb.n = d_out_n;
  cx_status st = run_search(h, ws, fh, sb, B, h->dim, ldq, kd, false, 0.0f, (uint32_t*)ws->hp, nullptr);
  // results are complete (run_search synchronised its stream); make the caller's stream see them
  cudaEventRecord(ev, ws->stream);
  cudaStreamWaitEvent(user, ev, 0);
  cudaEventDestroy(ev);
  return st;
}

extern "C" cx_status cx_row_id(const cx_index* h, uint32_t row, uint8_t out_id[16]) {
  if (!h || !out_id) return fail(CX_ERR_VALIDATION, "null argument");
  if (row >= h->n_rows) return fail(CX_ERR_VALIDATION, "row %u out of range", row);
  memcpy(out_id, &h->h_ids[(size_t)row * 16], 16);
  return CX_OK;
}

// ------------------------------------------------------------------------------
// save / load: bincode 1.3 (fixint, little endian) of
// (&HashMap<Uuid,Vec<f32>>, &HashMap<Uuid,NodeMetadata>, usize), index.rs:437-472.

static bool w64(FILE* fp, uint64_t v) { return fwrite(&v, 8, 1, fp) == 1; }
static bool wstr(FILE* fp, const std::string& s) {
  return w64(fp, s.size()) && (s.empty() || fwrite(s.data(), 1, s.size(), fp) == s.size());
}

extern "C" cx_status cx_save(const cx_index* h, const char* path) {
  if (!h || !path) return fail(CX_ERR_VALIDATION, "null argument");
  CU(cudaSetDevice(h->device));
  FILE* fp = fopen(path, "wb");
  if (!fp) return fail(CX_ERR_IO, "Failed to write index file: %s", path);
  bool ok = w64(fp, h->n_live);
  const uint64_t chunk = 65536;
  std::vector<float> buf((size_t)chunk * h->ld);
  for (uint64_t r0 = 0; r0 < h->n_rows && ok; r0 += chunk) {
    uint64_t n = h->n_rows - r0 < chunk ? h->n_rows - r0 : chunk;
    if (cudaMemcpy(buf.data(), h->dE + r0 * h->ld, n * h->ld * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
      fclose(fp);
      return fail(CX_ERR_CUDA, "device read failed during save");
    }
    for (uint64_t i = 0; i < n && ok; ++i) {
      uint32_t r = (uint32_t)(r0 + i);
      if (h->h_meta[r] & META_DEAD) continue;
      ok = w64(fp, 16) && fwrite(&h->h_ids[(size_t)r * 16], 16, 1, fp) == 1 && w64(fp, h->dim) &&
           fwrite(&buf[i * h->ld], 4, h->dim, fp) == h->dim;
    }
  }
  uint64_t n_meta = h->orphan_meta.size();
  for (uint32_t r = 0; r < h->n_rows; ++r)
    if (!(h->h_meta[r] & META_DEAD) && (h->h_meta[r] & META_HAS)) ++n_meta;
  ok = ok && w64(fp, n_meta);
  for (uint32_t r = 0; r < h->n_rows && ok; ++r) {
    uint32_t m = h->h_meta[r];
    if ((m & META_DEAD) || !(m & META_HAS)) continue;
    ok = w64(fp, 16) && fwrite(&h->h_ids[(size_t)r * 16], 16, 1, fp) == 1 &&
         wstr(fp, h->kinds.strs[m & META_KIND_MASK]) && wstr(fp, h->agents.strs[h->h_agent[r]]);
  }
  for (auto& kv : h->orphan_meta) {
    if (!ok) break;
    uint8_t id[16];
    memcpy(id, &kv.first.a, 8);
    memcpy(id + 8, &kv.first.b, 8);
    ok = w64(fp, 16) && fwrite(id, 16, 1, fp) == 1 && wstr(fp, h->kinds.strs[kv.second.first]) &&
         wstr(fp, h->agents.strs[kv.second.second]);
  }
  ok = ok && w64(fp, h->dim);
  ok = (fclose(fp) == 0) && ok;
  if (!ok) return fail(CX_ERR_IO, "Failed to write index file: %s", path);
  return CX_OK;
}

static bool r64(FILE* fp, uint64_t* v) { return fread(v, 8, 1, fp) == 1; }
static bool rstr(FILE* fp, std::string* s) {
  uint64_t n;
  if (!r64(fp, &n) || n > (1u << 24)) return false;
  s->resize(n);
  return n == 0 || fread(&(*s)[0], 1, n, fp) == n;
}

extern "C" cx_status cx_load(const char* path, int device, cx_index** out) {
  if (!path || !out) return fail(CX_ERR_VALIDATION, "null argument");
  *out = nullptr;
  FILE* fp = fopen(path, "rb");
  if (!fp) return fail(CX_ERR_IO, "Failed to read index file: %s", path);
  uint64_t dim = 0;
  if (fseek(fp, -8, SEEK_END) != 0 || !r64(fp, &dim) || dim == 0 || dim > 65536) {
    fclose(fp);
    return fail(CX_ERR_IO, "Failed to deserialize index: bad trailer in %s", path);
  }
  fseek(fp, 0, SEEK_SET);
  cx_index* h = nullptr;
  cx_status st = cx_index_create((uint32_t)dim, device, &h);
  if (st != CX_OK) {
    fclose(fp);
    return st;
  }
  auto bad = [&](const char* why) {
    fclose(fp);
    cx_index_destroy(h);
    return fail(CX_ERR_IO, "Failed to deserialize index: %s", why);
  };
  uint64_t n = 0, l = 0;
  if (!r64(fp, &n)) return bad("truncated header");
  const uint64_t chunk = 65536;
  std::vector<uint8_t> ids;
  std::vector<float> rows;
  for (uint64_t i0 = 0; i0 < n; i0 += chunk) {
    uint64_t m = n - i0 < chunk ? n - i0 : chunk;
    ids.resize(m * 16);
    rows.resize(m * dim);
    for (uint64_t i = 0; i < m; ++i) {
      if (!r64(fp, &l) || l != 16 || fread(&ids[i * 16], 16, 1, fp) != 1) return bad("bad id");
      if (!r64(fp, &l) || l != dim) return bad("vector length differs from dimension");
      if (fread(&rows[i * dim], 4, dim, fp) != dim) return bad("truncated vector");
    }
    st = cx_insert_batch(h, ids.data(), rows.data(), m, (uint32_t)dim);
    if (st != CX_OK) {
      fclose(fp);
      cx_index_destroy(h);
      return st;
    }
  }
  if (!r64(fp, &n)) return bad("truncated metadata header");
  for (uint64_t i = 0; i < n; ++i) {
    uint8_t id[16];
    std::string kind, agent;
    if (!r64(fp, &l) || l != 16 || fread(id, 16, 1, fp) != 1 || !rstr(fp, &kind) || !rstr(fp, &agent))
      return bad("bad metadata entry");
    st = cx_set_metadata(h, id, kind.c_str(), agent.c_str());
    if (st != CX_OK) {
      fclose(fp);
      cx_index_destroy(h);
      return st;
    }
  }
  fclose(fp);
  *out = h;
  return CX_OK;
}

// ------------------------------------------------------------------------------
extern "C" cx_status cx_get_stats(const cx_index* h, cx_stats* out) {
  if (!h || !out) return fail(CX_ERR_VALIDATION, "null argument");
  out->kernel_launches = h->launches.load();
  out->queries_stream = h->q_stream.load();
  out->queries_tensor = h->q_tensor.load();
  out->queries_exact = h->q_exact.load();
  out->fallbacks = h->fallbacks.load();
  out->h2d_bytes = h->h2d.load();
  out->d2h_bytes = h->d2h.load();
  out->pass_kernel_ns = h->pass_ns.load();
  out->pass_kernel_launches = h->pass_launches.load();
  return CX_OK;
}

extern "C" cx_status cx_set_option(cx_index* h, const char* key, int64_t value) {
  if (!h || !key) return fail(CX_ERR_VALIDATION, "null argument");
  if (!strcmp(key, "force_path")) {
    if (value < 0 || value > 3) return fail(CX_ERR_VALIDATION, "force_path must be 0..3");
    h->force_path = (int)value;
    return CX_OK;
  }
  if (!strcmp(key, "profile")) {
    h->profile = value != 0;
    return CX_OK;
  }
  if (!strcmp(key, "stream_max_batch")) {
    h->stream_max_batch = (uint32_t)value;
    return CX_OK;
  }
  return fail(CX_ERR_VALIDATION, "unknown option %s", key);
}

extern "C" const char* cx_last_error(void) { return g_err.c_str(); }
extern "C" const char* cx_version(void) { return "cortex_b200 0.1.0 (sm_100a)"; }
