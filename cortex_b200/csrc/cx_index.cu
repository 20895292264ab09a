// cx_index.cu -- the embedding store behind the C ABI (include/cortex_gpu.h):
// creation, insert / remove / set_metadata / rebuild, save / load, stats.
//
// Mirrors HnswIndex (vector/index.rs:182-473) in exact-scan mode:
//   vectors  : HashMap<NodeId, Vec<f32>>     -> device matrix + id->row hash on the host
//   metadata : HashMap<NodeId, NodeMetadata> -> per-row meta/agent words + string tables
// Data layout in HBM: E fp32 [rows][ld] row-major with 16 B aligned
// rows, a bf16 shadow [rows][ld16] for the tensor pass, norms / reciprocal norms,
// meta + agent words and the 16-byte ids, all indexed by row = insertion order.
// There is no CPU compute path: every score comes from a kernel.
#include <memory>

#include "cx_index.h"

using namespace cx;

// ------------------------------------------------------------------------------
static thread_local std::string g_err;

cx_status cx::fail(cx_status st, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return st;
}
const char* cx::last_error() { return g_err.c_str(); }

cudaError_t Workspace::ensure(size_t db, size_t hb) {
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

cudaError_t Workspace::ensure_state(size_t nq) {
  if (nq > state_q) {
    if (d_cnt) cudaFree(d_cnt);
    if (d_gtau) cudaFree(d_gtau);
    d_cnt = nullptr;
    d_gtau = nullptr;
    state_q = 0;
    size_t want = nq + nq / 2 + 64;
    cudaError_t e = cudaMalloc((void**)&d_cnt, want * sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    e = cudaMalloc((void**)&d_gtau, want * sizeof(uint64_t));
    if (e != cudaSuccess) return e;
    state_q = want;
    state_dirty = true;
  }
  if (state_dirty) {
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, state_q * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_gtau, 0, state_q * sizeof(uint64_t), stream);
    if (e != cudaSuccess) return e;
    state_dirty = false;
  }
  return cudaSuccess;
}

Workspace::~Workspace() {
  if (d) cudaFree(d);
  if (hp) cudaFreeHost(hp);
  if (d_cnt) cudaFree(d_cnt);
  if (d_gtau) cudaFree(d_gtau);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (ev_sync) cudaEventDestroy(ev_sync);
  if (ev_block) cudaEventDestroy(ev_block);
  if (stream) cudaStreamDestroy(stream);
}

cudaError_t WsLease::init() {
  if (ws) return cudaSuccess;
  ws = new Workspace();
  cudaError_t e = cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) return e;
  e = cudaEventCreate(&ws->ev0);
  if (e != cudaSuccess) return e;
  e = cudaEventCreate(&ws->ev1);
  if (e != cudaSuccess) return e;
  e = cudaEventCreateWithFlags(&ws->ev_block, cudaEventDisableTiming | cudaEventBlockingSync);
  if (e != cudaSuccess) return e;
  return cudaEventCreateWithFlags(&ws->ev_sync, cudaEventDisableTiming);
}

// Wait for the workspace's stream.  blocking: sleep on an event instead of spinning, which leaves
// the core to other threads (one process per GPU on a host with few cores per GPU).
cudaError_t Workspace::wait(bool blocking) {
  if (!blocking) return cudaStreamSynchronize(stream);
  cudaError_t e = cudaEventRecord(ev_block, stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev_block);
}

// Count the irregular rows (NaN reciprocal norm, cx_exact.cu) of the whole store into h->n_irr.
// Enqueued on s; the value is valid after the caller's stream synchronisation.
static cudaError_t refresh_irregular(cx_index* h, cudaStream_t s) {
  if (!h->d_irr) {
    cudaError_t e = cudaMalloc((void**)&h->d_irr, 4);
    if (e != cudaSuccess) return e;
    e = cudaMallocHost((void**)&h->h_irr, 4);
    if (e != cudaSuccess) return e;
    *h->h_irr = 0;
  }
  launch_count_irregular(h->dRnorm, (uint32_t)h->n_rows, h->d_irr, s);
  return cudaMemcpyAsync(h->h_irr, h->d_irr, 4, cudaMemcpyDeviceToHost, s);
}

// ------------------------------------------------------------------------------
static void free_store(cx_index* h) {
  cudaFree(h->dE);
  cudaFree(h->dNorm);
  cudaFree(h->dRnorm);
  cudaFree(h->dMeta);
  cudaFree(h->dAgent);
  cudaFree(h->dIds);
  if (h->dE16) cudaFree(h->dE16);
  h->dE = h->dNorm = h->dRnorm = nullptr;
  h->dMeta = h->dAgent = nullptr;
  h->dIds = nullptr;
  h->dE16 = nullptr;
}

struct StoreAlloc {
  float *E = nullptr, *nm = nullptr, *rn = nullptr;
  uint32_t *me = nullptr, *ag = nullptr;
  uint8_t* ids = nullptr;
  void* e16 = nullptr;
};

static cx_status alloc_store(const cx_index* h, uint64_t ncap, StoreAlloc* a) {
  CU(cudaMalloc(&a->E, ncap * h->ld * sizeof(float)));
  // +32: the streaming pass fetches reciprocal norms in 16 B units past the last row
  CU(cudaMalloc(&a->nm, (ncap + 32) * sizeof(float)));
  CU(cudaMalloc(&a->rn, (ncap + 32) * sizeof(float)));
  CU(cudaMalloc(&a->me, ncap * sizeof(uint32_t)));
  CU(cudaMalloc(&a->ag, ncap * sizeof(uint32_t)));
  CU(cudaMalloc(&a->ids, ncap * 16));
  if (h->want_shadow) CU(cudaMalloc(&a->e16, ncap * h->ld16 * 2));
  return CX_OK;
}

static void adopt_store(cx_index* h, const StoreAlloc& a, uint64_t ncap) {
  free_store(h);
  h->dE = a.E;
  h->dNorm = a.nm;
  h->dRnorm = a.rn;
  h->dMeta = a.me;
  h->dAgent = a.ag;
  h->dIds = a.ids;
  h->dE16 = a.e16;
  h->cap = ncap;
}

static cx_status grow(cx_index* h, uint64_t need) {
  if (need <= h->cap) return CX_OK;
  if (need > 0x7FFFFF00ull) return fail(CX_ERR_VALIDATION, "index shard limited to 2^31 rows");
  uint64_t ncap = h->cap ? h->cap * 2 : 1024;
  if (ncap < need) ncap = need;
  StoreAlloc a;
  cx_status st = alloc_store(h, ncap, &a);
  if (st != CX_OK) return st;
  cudaStream_t s = h->mut_stream;
  if (h->n_rows) {
    CU(cudaMemcpyAsync(a.E, h->dE, h->n_rows * h->ld * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(a.nm, h->dNorm, h->n_rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(a.rn, h->dRnorm, h->n_rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(a.me, h->dMeta, h->n_rows * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(a.ag, h->dAgent, h->n_rows * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    CU(cudaMemcpyAsync(a.ids, h->dIds, h->n_rows * 16, cudaMemcpyDeviceToDevice, s));
    if (a.e16 && h->dE16)
      CU(cudaMemcpyAsync(a.e16, h->dE16, h->n_rows * h->ld16 * 2, cudaMemcpyDeviceToDevice, s));
  }
  CU(cudaStreamSynchronize(s));
  adopt_store(h, a, ncap);
  h->h_ids.reserve(ncap * 16);
  return CX_OK;
}

static uint32_t meta_word(bool has, uint32_t kind) { return has ? (META_HAS | (kind & META_KIND_MASK)) : 0u; }

extern "C" cx_status cx_index_create(uint32_t dimension, int device, cx_index** out) {
  if (!out) return fail(CX_ERR_VALIDATION, "out is null");
  *out = nullptr;
  if (dimension == 0 || dimension > 65536)
    return fail(CX_ERR_VALIDATION, "dimension %u out of range", dimension);
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
  free_store(h);
  if (h->d_irr) cudaFree(h->d_irr);
  if (h->h_irr) cudaFreeHost(h->h_irr);
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
    uint64_t j = i;  // run of new ids
    const uint64_t r0 = h->n_rows;
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
    CU(cudaMemcpy2DAsync(h->dE + (size_t)r0 * h->ld, h->ld * sizeof(float), rows + (size_t)i * len,
                         len * sizeof(float), len * sizeof(float), j - i, cudaMemcpyHostToDevice, s));
    i = j;
  }
  const uint64_t n_new = h->n_rows - first_new;
  h->h2d += n * len * sizeof(float);
  if (n_new) {
    CU(cudaMemcpyAsync(h->dIds + first_new * 16, h->h_ids.data() + first_new * 16, n_new * 16,
                       cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->dMeta + first_new, h->h_meta.data() + first_new, n_new * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->dAgent + first_new, h->h_agent.data() + first_new, n_new * 4, cudaMemcpyHostToDevice,
                       s));
    launch_prepare_rows(h->dE, h->dNorm, h->dRnorm, h->dE16, h->dim, h->ld, h->ld16, (uint32_t)first_new,
                        (uint32_t)n_new, s);
    h->launches += h->dE16 ? 2 : 1;
  }
  for (uint32_t r : overwritten) {
    launch_prepare_rows(h->dE, h->dNorm, h->dRnorm, h->dE16, h->dim, h->ld, h->ld16, r, 1, s);
    h->launches += h->dE16 ? 2 : 1;
  }
  CU(refresh_irregular(h, s));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s));
  return CX_OK;
}

// Bulk append of rows that already live in device memory (row-major [n][len] f32): the
// "mmap'd vectors bulk-uploaded once" path of the north star without a host round trip.
// All ids must be new (this is an append, not an upsert).
extern "C" cx_status cx_insert_batch_device(cx_index* h, const uint8_t* ids, const float* d_rows, uint64_t n,
                                            uint32_t len) {
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (len != h->dim)
    return fail(CX_ERR_VALIDATION, "Embedding dimension mismatch: expected %u, got %u", h->dim, len);
  if (!n) return CX_OK;
  if (!ids || !d_rows) return fail(CX_ERR_VALIDATION, "null ids/rows");
  for (uint64_t i = 0; i < n; ++i)
    if (h->id2row.count(load_id(ids + 16 * i)))
      return fail(CX_ERR_VALIDATION, "cx_insert_batch_device appends only: id %llu already present",
                  (unsigned long long)i);
  CU(cudaSetDevice(h->device));
  cx_status st = grow(h, h->n_rows + n);
  if (st != CX_OK) return st;
  cudaStream_t s = h->mut_stream;
  const uint64_t first_new = h->n_rows;
  h->h_ids.insert(h->h_ids.end(), ids, ids + 16 * n);
  for (uint64_t i = 0; i < n; ++i) {
    Id128 key = load_id(ids + 16 * i);
    if (!h->id2row.emplace(key, (uint32_t)(first_new + i)).second) {
      return fail(CX_ERR_VALIDATION, "duplicate id inside the batch at %llu", (unsigned long long)i);
    }
    uint32_t mw = 0, ag = 0;
    auto om = h->orphan_meta.find(key);
    if (om != h->orphan_meta.end()) {
      mw = meta_word(true, om->second.first);
      ag = om->second.second;
      h->orphan_meta.erase(om);
    }
    h->h_meta.push_back(mw);
    h->h_agent.push_back(ag);
  }
  h->n_rows += n;
  h->n_live += n;
  // the caller's buffer may have been produced on another stream: order after the device
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy2DAsync(h->dE + first_new * h->ld, h->ld * sizeof(float), d_rows, len * sizeof(float),
                       len * sizeof(float), n, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(h->dIds + first_new * 16, h->h_ids.data() + first_new * 16, n * 16, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->dMeta + first_new, h->h_meta.data() + first_new, n * 4, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->dAgent + first_new, h->h_agent.data() + first_new, n * 4, cudaMemcpyHostToDevice, s));
  launch_prepare_rows(h->dE, h->dNorm, h->dRnorm, h->dE16, h->dim, h->ld, h->ld16, (uint32_t)first_new, (uint32_t)n, s);
  h->launches += h->dE16 ? 2 : 1;
  CU(refresh_irregular(h, s));
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
  if (it == h->id2row.end()) {  // the metadata map is independent of the vector map (index.rs:219-222)
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
  const uint64_t ncap = nl > 1024 ? nl : 1024;
  StoreAlloc a;
  cx_status st = alloc_store(h, ncap, &a);
  if (st != CX_OK) return st;
  uint32_t* dlive = nullptr;
  if (nl) {
    CU(cudaMalloc(&dlive, nl * 4));
    CU(cudaMemcpyAsync(dlive, live.data(), nl * 4, cudaMemcpyHostToDevice, s));
    launch_gather_rows(h->view(), a.E, a.nm, a.rn, a.me, a.ag, a.ids, a.e16, dlive, (uint32_t)nl, s);
    h->launches += 1;
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(s));
  if (dlive) cudaFree(dlive);
  adopt_store(h, a, ncap);
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
  CU(refresh_irregular(h, s));  // removed rows are gone: recount
  CU(cudaStreamSynchronize(s));
  return CX_OK;
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
  if (!strcmp(key, "tensor_pair")) {  // 0 = single-CTA tcgen05 form only (A/B measurements); process-wide
    tensor_set_pair(value != 0);
    return CX_OK;
  }
  if (!strcmp(key, "tensor_phase_growth")) {
    if (value < 0 || value > 1024) return fail(CX_ERR_VALIDATION, "tensor_phase_growth must be 0..1024");
    h->tensor_phase_growth = (uint32_t)value;
    return CX_OK;
  }
  if (!strcmp(key, "blocking_sync")) {
    h->blocking_sync = value != 0;
    return CX_OK;
  }
  if (!strcmp(key, "tensor_epi_warps")) {
    tensor_set_epi_warps((int)value);
    return CX_OK;
  }
  if (!strcmp(key, "tensor_debug")) {  // measurement hook: results of the tensor pass become wrong
    tensor_set_debug((int)value);
    return CX_OK;
  }
  if (!strcmp(key, "tensor_min_batch")) {
    if (value < 1) return fail(CX_ERR_VALIDATION, "tensor_min_batch must be >= 1");
    h->tensor_min_batch = (uint32_t)value;
    return CX_OK;
  }
  if (!strcmp(key, "shadow")) {  // bf16 shadow matrix for the tensor pass; only before the first insert
    if (h->n_rows) return fail(CX_ERR_VALIDATION, "shadow can only be changed on an empty index");
    h->want_shadow = value != 0;
    return CX_OK;
  }
  return fail(CX_ERR_VALIDATION, "unknown option %s", key);
}

extern "C" const char* cx_last_error(void) { return cx::last_error(); }
extern "C" const char* cx_version(void) { return "cortex_b200 0.1.0 (sm_100a)"; }
