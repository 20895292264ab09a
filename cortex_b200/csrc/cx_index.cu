// cx_index.cu -- the embedding store behind the C ABI (include/cortex_gpu.h):
// creation, insert / remove / set_metadata / rebuild, save / load, stats.
//
// Mirrors HnswIndex (vector/index.rs:182-473) in exact-scan mode:
//   vectors  : HashMap<NodeId, Vec<f32>>     -> device matrix + id->row hash on the host
//   metadata : HashMap<NodeId, NodeMetadata> -> per-row meta/agent words + string tables
// Data layout in HBM: E fp32 [rows][ld] row-major with 16 B aligned
// rows, a bf16 shadow [rows][ld16] for the tensor pass, norms / reciprocal norms,
// meta + agent words and the 16-byte ids, all indexed by row = insertion order.
// Every array grows IN PLACE behind a reserved range of device addresses (VmArray): appending
// rows never copies or reallocates the store, so ingest latency has no growth spikes and the
// store never exists twice.  Mutations are enqueued on the index's own stream and return
// without waiting; the next search (or mutation) waits for them once (cx::settle).
// There is no CPU compute path: every score comes from a kernel.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
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

// ---- driver virtual-memory API (resolved at run time; libcuda is not linked) -----------------
namespace {
struct VmApi {
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
  CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};

template <typename F>
bool resolve(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    (void)cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

const VmApi& vm_api() {
  static VmApi api = [] {
    VmApi a;
    const char* off = getenv("CORTEX_GPU_NO_VMM");  // test hook: exercise the allocate-copy-free fallback
    if (off && off[0] == '1') return a;
    a.ok = resolve("cuMemAddressReserve", &a.reserve) && resolve("cuMemCreate", &a.create) &&
           resolve("cuMemMap", &a.map) && resolve("cuMemSetAccess", &a.set_access) &&
           resolve("cuMemUnmap", &a.unmap) && resolve("cuMemRelease", &a.release) &&
           resolve("cuMemAddressFree", &a.addr_free) &&
           resolve("cuMemGetAllocationGranularity", &a.granularity);
    return a;
  }();
  return api;
}

CUmemAllocationProp vm_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  return prop;
}

size_t vm_granularity(int device) {
  size_t g = 0;
  const CUmemAllocationProp prop = vm_prop(device);
  if (vm_api().granularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !g) g = 2u << 20;
  return g;
}
}  // namespace

bool cx::vmm_available() { return vm_api().ok; }

cx_status VmArray::init(int dev, size_t reserve_bytes) {
  device = dev;
  p = nullptr;
  mapped = 0;
  reserved = 0;
  vmm = false;
  chunks.clear();
  if (!vm_api().ok) return CX_OK;  // fallback arrays are allocated on first use
  const size_t g = vm_granularity(dev);
  const size_t want = align_up(reserve_bytes ? reserve_bytes : g, g);
  CUdeviceptr base = 0;
  if (vm_api().reserve(&base, want, 0, 0, 0) != CUDA_SUCCESS) return CX_OK;  // no address space: fallback
  p = reinterpret_cast<char*>(base);
  reserved = want;
  vmm = true;
  return CX_OK;
}

cx_status VmArray::ensure(size_t bytes, size_t used, cudaStream_t s) {
  if (bytes <= mapped) return CX_OK;
  if (vmm) {
    const size_t g = vm_granularity(device);
    const size_t upto = align_up(bytes, g);
    if (upto > reserved)
      return fail(CX_ERR_CUDA, "embedding store: %zu bytes exceed the %zu reserved for this device", upto, reserved);
    // slices of at most 256 MB: the driver holds its lock while it maps, and kernel launches of
    // concurrent searches wait for that lock -- short holds keep them flowing
    const size_t slice = align_up((size_t)256 << 20, g);
    const CUmemAllocationProp prop = vm_prop(device);
    while (mapped < upto) {
      const size_t add = upto - mapped < slice ? upto - mapped : slice;
      CUmemGenericAllocationHandle hd = 0;
      CUresult r = vm_api().create(&hd, add, &prop, 0);
      if (r != CUDA_SUCCESS)
        return fail(CX_ERR_CUDA, "cuMemCreate(%zu bytes) failed (%d): out of device memory", add, (int)r);
      const CUdeviceptr at = reinterpret_cast<CUdeviceptr>(p) + mapped;
      r = vm_api().map(at, add, 0, hd, 0);
      if (r != CUDA_SUCCESS) {
        vm_api().release(hd);
        return fail(CX_ERR_CUDA, "cuMemMap failed (%d)", (int)r);
      }
      CUmemAccessDesc acc;
      memset(&acc, 0, sizeof acc);
      acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      acc.location.id = device;
      acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      r = vm_api().set_access(at, add, &acc, 1);
      if (r != CUDA_SUCCESS) {
        vm_api().unmap(at, add);
        vm_api().release(hd);
        return fail(CX_ERR_CUDA, "cuMemSetAccess failed (%d)", (int)r);
      }
      chunks.push_back({(unsigned long long)hd, mapped, add});
      mapped += add;
    }
    return CX_OK;
  }
  // fallback: allocate, copy what is in use, free
  char* np = nullptr;
  CU(cudaMalloc((void**)&np, bytes));
  if (p && used) {
    cudaError_t e = cudaMemcpyAsync(np, p, used, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      cudaFree(np);
      return fail(CX_ERR_CUDA, "CUDA error %s while growing the store", cudaGetErrorString(e));
    }
  }
  if (p) cudaFree(p);
  p = np;
  mapped = bytes;
  return CX_OK;
}

void VmArray::shrink(size_t bytes) {
  if (!vmm) return;
  const size_t keep = align_up(bytes, vm_granularity(device));
  while (!chunks.empty() && chunks.back().off >= keep) {
    const Chunk c = chunks.back();
    chunks.pop_back();
    vm_api().unmap(reinterpret_cast<CUdeviceptr>(p) + c.off, c.size);
    vm_api().release((CUmemGenericAllocationHandle)c.handle);
    mapped = c.off;
  }
}

void VmArray::destroy() {
  if (vmm) {
    for (const Chunk& c : chunks) {
      vm_api().unmap(reinterpret_cast<CUdeviceptr>(p) + c.off, c.size);
      vm_api().release((CUmemGenericAllocationHandle)c.handle);
    }
    if (p) vm_api().addr_free(reinterpret_cast<CUdeviceptr>(p), reserved);
  } else if (p) {
    cudaFree(p);
  }
  chunks.clear();
  p = nullptr;
  mapped = reserved = 0;
  vmm = false;
}

// ---- per-call workspaces -----------------------------------------------------------------
static cudaError_t regrow(void** p, size_t* have, size_t want, bool pinned) {
  if (want <= *have) return cudaSuccess;
  if (*p) {
    if (pinned) cudaFreeHost(*p);
    else cudaFree(*p);
  }
  *p = nullptr;
  *have = 0;
  const size_t sz = want + want / 4;
  cudaError_t e = pinned ? cudaMallocHost(p, sz) : cudaMalloc(p, sz);
  if (e != cudaSuccess) return e;
  *have = sz;
  return cudaSuccess;
}

cudaError_t Workspace::ensure(size_t db, size_t hb) {
  if (db > d_bytes || hb > h_bytes) drop_graph();  // a recorded launch sequence points into these buffers
  cudaError_t e = regrow(&d, &d_bytes, db, false);
  if (e != cudaSuccess) return e;
  return regrow(&hp, &h_bytes, hb, true);
}

cudaError_t Workspace::ensure_exact(size_t bytes) { return regrow(&dx, &dx_bytes, bytes, false); }

cudaError_t Workspace::ensure_aux(size_t db, size_t hb) {
  cudaError_t e = regrow(&aux, &aux_bytes, db, false);
  if (e != cudaSuccess) return e;
  return regrow(&aux_h, &aux_h_bytes, hb, true);
}

void Workspace::drop_graph() {
  if (graph) cudaGraphExecDestroy(graph);
  graph = nullptr;
  memset(graph_key, 0, sizeof graph_key);
  memset(last_key, 0, sizeof last_key);
}

cudaError_t Workspace::ensure_state(size_t nq) {
  if (nq > state_q) {
    drop_graph();
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
  if (graph) cudaGraphExecDestroy(graph);
  if (d) cudaFree(d);
  if (hp) cudaFreeHost(hp);
  if (dx) cudaFree(dx);
  if (aux) cudaFree(aux);
  if (aux_h) cudaFreeHost(aux_h);
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
  std::unique_ptr<Workspace> w(new Workspace());  // a half-built workspace is destroyed, never pooled
  cudaError_t e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) return e;
  e = cudaEventCreate(&w->ev0);
  if (e != cudaSuccess) return e;
  e = cudaEventCreate(&w->ev1);
  if (e != cudaSuccess) return e;
  e = cudaEventCreateWithFlags(&w->ev_block, cudaEventDisableTiming | cudaEventBlockingSync);
  if (e != cudaSuccess) return e;
  e = cudaEventCreateWithFlags(&w->ev_sync, cudaEventDisableTiming);
  if (e != cudaSuccess) return e;
  ws = w.release();
  return cudaSuccess;
}

// Wait for the workspace's stream.  blocking: sleep on an event instead of spinning, which leaves
// the core to other threads (one process per GPU on a host with few cores per GPU).
cudaError_t Workspace::wait(bool blocking) {
  if (!blocking) return cudaStreamSynchronize(stream);
  cudaError_t e = cudaEventRecord(ev_block, stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev_block);
}

// ---- mutation ordering ---------------------------------------------------------------------
cx_status cx::settle(cx_index* h) {
  if (h->mut_dirty.load(std::memory_order_acquire)) {
    CU(cudaEventSynchronize(h->ev_mut));
    h->mut_dirty.store(false, std::memory_order_release);
  }
  return CX_OK;
}

// every mutation ends here: publish the irregular-row count and mark the stream position
static cx_status mutation_done(cx_index* h) {
  cudaStream_t s = h->mut_stream;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h->h_irr, h->d_irr, 4, cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(h->ev_mut, s));
  h->mut_dirty.store(true, std::memory_order_release);
  return CX_OK;
}

// pinned staging for the small per-row side arrays of an insert (ids, meta, agent, seq)
static cx_status stage(cx_index* h, size_t bytes) {
  if (bytes <= h->stage_bytes) return CX_OK;
  if (h->stage_h) cudaFreeHost(h->stage_h);
  h->stage_h = nullptr;
  h->stage_bytes = 0;
  const size_t want = bytes + bytes / 2 + 4096;
  CU(cudaMallocHost(&h->stage_h, want));
  h->stage_bytes = want;
  return CX_OK;
}

// ---- the store ---------------------------------------------------------------------------------
static void refresh_ptrs(cx_index* h) {
  h->dE = (float*)h->aE.p;
  h->dNorm = (float*)h->aNorm.p;
  h->dRnorm = (float*)h->aRnorm.p;
  h->dMeta = (uint32_t*)h->aMeta.p;
  h->dAgent = (uint32_t*)h->aAgent.p;
  h->dIds = (uint8_t*)h->aIds.p;
  h->dE16 = h->want_shadow ? (void*)h->aE16.p : nullptr;
  h->dSeq = h->with_seq ? (uint64_t*)h->aSeq.p : nullptr;
}

static void free_store(cx_index* h) {
  h->aE.destroy();
  h->aNorm.destroy();
  h->aRnorm.destroy();
  h->aMeta.destroy();
  h->aAgent.destroy();
  h->aIds.destroy();
  h->aE16.destroy();
  h->aSeq.destroy();
  refresh_ptrs(h);
  h->cap = 0;
}

// address ranges for the largest shard this device could ever hold (its whole memory in rows)
static cx_status reserve_store(cx_index* h, size_t total_mem) {
  const size_t row_bytes = (size_t)h->ld * 4 + (h->want_shadow ? (size_t)h->ld16 * 2 : 0) + 4 + 4 + 4 + 4 + 16 +
                           (h->with_seq ? 8 : 0);
  uint64_t mr = total_mem / row_bytes + 65536;
  if (mr > 0x7FFFFF00ull) mr = 0x7FFFFF00ull;
  h->max_rows = mr;
  cx_status st;
  if ((st = h->aE.init(h->device, mr * h->ld * 4)) != CX_OK) return st;
  // +32: the streaming pass fetches reciprocal norms in 16 B units past the last row
  if ((st = h->aNorm.init(h->device, (mr + 32) * 4)) != CX_OK) return st;
  if ((st = h->aRnorm.init(h->device, (mr + 32) * 4)) != CX_OK) return st;
  if ((st = h->aMeta.init(h->device, mr * 4)) != CX_OK) return st;
  if ((st = h->aAgent.init(h->device, mr * 4)) != CX_OK) return st;
  if ((st = h->aIds.init(h->device, mr * 16)) != CX_OK) return st;
  if ((st = h->aE16.init(h->device, h->want_shadow ? mr * h->ld16 * 2 : 0)) != CX_OK) return st;
  if ((st = h->aSeq.init(h->device, h->with_seq ? mr * 8 : 0)) != CX_OK) return st;
  return CX_OK;
}

// map every array up to ncap rows (no copy in vmm mode; the fallback copies the rows in use on s)
static cx_status map_store(cx_index* h, uint64_t ncap, cudaStream_t s) {
  const uint64_t u = h->n_rows;
  cx_status st;
  if ((st = h->aE.ensure(ncap * h->ld * 4, u * h->ld * 4, s)) != CX_OK) return st;
  if ((st = h->aNorm.ensure((ncap + 32) * 4, u * 4, s)) != CX_OK) return st;
  if ((st = h->aRnorm.ensure((ncap + 32) * 4, u * 4, s)) != CX_OK) return st;
  if ((st = h->aMeta.ensure(ncap * 4, u * 4, s)) != CX_OK) return st;
  if ((st = h->aAgent.ensure(ncap * 4, u * 4, s)) != CX_OK) return st;
  if ((st = h->aIds.ensure(ncap * 16, u * 16, s)) != CX_OK) return st;
  if (h->want_shadow && (st = h->aE16.ensure(ncap * h->ld16 * 2, u * h->ld16 * 2, s)) != CX_OK) return st;
  if (h->with_seq && (st = h->aSeq.ensure(ncap * 8, u * 8, s)) != CX_OK) return st;
  return CX_OK;
}

static void join_grower(cx_index* h) {
  if (h->grow_thread.joinable()) h->grow_thread.join();
}

static uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

static cx_status grow(cx_index* h, uint64_t need, bool exact = false) {
  if (need <= h->cap.load()) return CX_OK;
  if (need > 0x7FFFFF00ull) return fail(CX_ERR_VALIDATION, "index shard limited to 2^31 rows");
  if (h->grow_thread.joinable()) {  // an extension is already on its way: it normally covers the need
    h->grow_waits += 1;
    join_grower(h);
    if (need <= h->cap.load()) return CX_OK;
  }
  const bool vmm = h->aE.vmm;
  const uint64_t cap = h->cap.load();
  uint64_t ncap = need;
  if (!exact) {
    // in-place growth is cheap: a quarter more at a time bounds the unused tail to 25 %; the
    // copying fallback doubles
    const uint64_t step = vmm ? (cap / 4 > 4096 ? cap / 4 : 4096) : (cap > 1024 ? cap : 1024);
    if (cap + step > ncap) ncap = cap + step;
  }
  if (vmm && ncap > h->max_rows) ncap = need;
  const uint64_t t0 = now_ns();
  cx_status st = map_store(h, ncap, h->mut_stream);
  if (st != CX_OK) return st;
  const uint64_t dt = now_ns() - t0;
  h->grow_ns += dt;
  if (dt > h->grow_ns_max.load()) h->grow_ns_max = dt;
  refresh_ptrs(h);
  h->cap = ncap;
  h->grow_events += 1;
  // pointers may have moved (fallback arrays): recorded launch sequences are stale
  if (!vmm) {
    std::lock_guard<std::mutex> g(h->ws_mu);
    for (Workspace* w : h->ws_free) w->drop_graph();
  }
  return CX_OK;
}

// Called at the end of an insert: when the mapped store is nearly used up, map the next extension on a helper
// thread.  Only in vmm mode (pointers never move; mapping new pages does not touch pages in use).
static void grow_ahead(cx_index* h, uint64_t n_new) {
  if (!h->aE.vmm || h->grow_thread.joinable()) return;
  const uint64_t cap = h->cap.load();
  if (h->n_rows + cap / 8 + 1024 < cap) return;  // more than an eighth of the capacity is still free
  // only streaming ingest is worth running ahead of: a bulk load that just filled what the caller reserved
  // says nothing about what comes next, and mapping gigabytes behind its back would compete with the
  // searches that follow
  if (n_new > cap / 64 + 4096) return;
  uint64_t ncap = cap + (cap / 4 > 4096 ? cap / 4 : 4096);
  if (ncap > h->max_rows) ncap = h->max_rows;
  if (ncap <= cap) return;
  h->grow_thread = std::thread([h, ncap] {
    cudaSetDevice(h->device);
    const uint64_t t0 = now_ns();
    const cx_status st = map_store(h, ncap, nullptr);
    const uint64_t dt = now_ns() - t0;
    h->grow_status = st;
    if (st != CX_OK) {
      h->grow_error = last_error();
      return;  // the next insert that needs the room maps it itself and reports the error
    }
    h->grow_ns += dt;
    if (dt > h->grow_ns_max.load()) h->grow_ns_max = dt;
    h->grow_events += 1;
    h->cap = ncap;
  });
}

static uint32_t meta_word(bool has, uint32_t kind) { return has ? (META_HAS | (kind & META_KIND_MASK)) : 0u; }

cx_status cx::index_create(uint32_t dimension, int device, bool with_seq, cx_index** out) {
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
  CU(cudaFree(0));  // make sure the primary context exists before driver-level calls
  std::unique_ptr<cx_index, void (*)(cx_index*)> h(new cx_index(), cx_index_destroy);
  h->device = device;
  h->dim = dimension;
  h->ld = (uint32_t)align_up(dimension, 4);
  h->ld16 = (uint32_t)align_up(dimension, 64);
  h->with_seq = with_seq;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(CX_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                prop.minor);
  h->sm_count = prop.multiProcessorCount;
  h->total_mem = prop.totalGlobalMem;
  CU(cudaStreamCreateWithFlags(&h->mut_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&h->ev_mut, cudaEventDisableTiming));
  CU(cudaMalloc((void**)&h->d_irr, 4));
  CU(cudaMemset(h->d_irr, 0, 4));
  CU(cudaMallocHost((void**)&h->h_irr, 4));
  *h->h_irr = 0;
  *out = h.release();
  return CX_OK;
}

extern "C" cx_status cx_index_create(uint32_t dimension, int device, cx_index** out) {
  cx::CallerDevice keep_callers_device;
  if (!out) return fail(CX_ERR_VALIDATION, "out is null");
  return index_create(dimension, device, false, out);
}

extern "C" void cx_index_destroy(cx_index* h) {
  cx::CallerDevice keep_callers_device;
  if (!h) return;
  if (h->shards) {
    shard_destroy(h);
    delete h;
    return;
  }
  cudaSetDevice(h->device);
  join_grower(h);
  if (h->mut_stream) cudaStreamSynchronize(h->mut_stream);
  for (Workspace* w : h->ws_free) delete w;
  free_store(h);
  if (h->d_irr) cudaFree(h->d_irr);
  if (h->h_irr) cudaFreeHost(h->h_irr);
  if (h->stage_h) cudaFreeHost(h->stage_h);
  if (h->ev_mut) cudaEventDestroy(h->ev_mut);
  if (h->mut_stream) cudaStreamDestroy(h->mut_stream);
  delete h;
}

extern "C" cx_status cx_reserve(cx_index* h, uint64_t n_rows) {
  cx::CallerDevice keep_callers_device;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (h->shards) return shard_reserve(h, n_rows);
  CU(cudaSetDevice(h->device));
  cx_status st = settle(h);
  if (st != CX_OK) return st;
  if (!h->store_ready && (st = index_prepare_store(h)) != CX_OK) return st;
  join_grower(h);
  return grow(h, n_rows, /*exact=*/true);
}

cx_status cx::index_prepare_store(cx_index* h) {
  if (h->store_ready) return CX_OK;
  cx_status st = reserve_store(h, h->total_mem);
  if (st != CX_OK) return st;
  h->store_ready = true;
  return CX_OK;
}

// Upsert of n rows (HashMap::insert, index.rs:307: the same id overwrites in place).  `rows` is
// row-major [n][len] in host memory or -- on_device -- in this device's memory.  seq (shards of a
// multi-device index): global insertion number of each row.  Nothing of the host-side state changes
// unless every device operation could be enqueued.
cx_status cx::index_insert(cx_index* h, const uint8_t* ids, const float* rows, uint64_t n, uint32_t len,
                           bool on_device, const uint64_t* seq) {
  if (len != h->dim)
    return fail(CX_ERR_VALIDATION, "Embedding dimension mismatch: expected %u, got %u", h->dim, len);
  if (!n) return CX_OK;
  if (!ids || !rows) return fail(CX_ERR_VALIDATION, "null ids/rows");
  if (h->with_seq && !seq) return fail(CX_ERR_VALIDATION, "shard insert without sequence numbers");
  CU(cudaSetDevice(h->device));
  cx_status st = settle(h);  // the previous mutation has finished: staging buffers are free again
  if (st != CX_OK) return st;
  if (!h->store_ready && (st = index_prepare_store(h)) != CX_OK) return st;
  cudaStream_t s = h->mut_stream;

  // ---- 1. where does every row go?  (no state is touched yet)
  const uint64_t first_new = h->n_rows;
  std::vector<uint32_t> tgt(n);
  std::vector<uint32_t> overwritten;
  uint64_t n_new = 0;
  {
    std::unordered_map<Id128, uint32_t, Id128Hash> fresh;  // ids first seen in this batch
    if (n > 1) fresh.reserve(n);
    for (uint64_t i = 0; i < n; ++i) {
      const Id128 key = load_id(ids + 16 * i);
      if (const uint32_t* row = h->id2row.find(key)) {
        tgt[i] = *row;
        overwritten.push_back(*row);
        continue;
      }
      if (n > 1) {
        auto f = fresh.find(key);
        if (f != fresh.end()) {  // the same new id twice in one batch: the later row wins
          tgt[i] = f->second;
          continue;
        }
        fresh.emplace(key, (uint32_t)(first_new + n_new));
      }
      tgt[i] = (uint32_t)(first_new + n_new++);
    }
  }
  std::sort(overwritten.begin(), overwritten.end());
  overwritten.erase(std::unique(overwritten.begin(), overwritten.end()), overwritten.end());
  st = grow(h, first_new + n_new);
  if (st != CX_OK) return st;

  // ---- 2. enqueue the device work
  // rows that are about to be overwritten leave the irregular count with their old contents
  for (uint32_t r : overwritten) launch_count_irregular(h->dRnorm, h->dMeta, r, 1, -1, h->d_irr, s);
  if (on_device) CU(cudaDeviceSynchronize());  // the caller's buffer may have been produced on another stream
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (uint64_t i = 0; i < n;) {
    uint64_t j = i + 1;
    while (j < n && tgt[j] == tgt[j - 1] + 1) ++j;  // run of consecutive target rows: one strided copy
    CU(cudaMemcpy2DAsync(h->dE + (size_t)tgt[i] * h->ld, h->ld * sizeof(float), rows + (size_t)i * len,
                         len * sizeof(float), len * sizeof(float), j - i, kind, s));
    i = j;
  }
  if (!on_device) h->h2d += n * len * sizeof(float);
  // side arrays of the new rows through pinned staging: [ids | meta | agent | seq]
  uint8_t* s_ids = nullptr;
  uint32_t *s_meta = nullptr, *s_agent = nullptr;
  uint64_t* s_seq = nullptr;
  if (n_new) {
    const size_t o_seq = 0, o_ids = o_seq + n_new * 8, o_meta = o_ids + n_new * 16, o_agent = o_meta + n_new * 4;
    st = stage(h, o_agent + n_new * 4);
    if (st != CX_OK) return st;
    char* sp = (char*)h->stage_h;
    s_seq = (uint64_t*)(sp + o_seq);
    s_ids = (uint8_t*)(sp + o_ids);
    s_meta = (uint32_t*)(sp + o_meta);
    s_agent = (uint32_t*)(sp + o_agent);
    uint64_t seen = 0;
    for (uint64_t i = 0; i < n; ++i) {
      if (tgt[i] < first_new) continue;
      const uint64_t o = tgt[i] - first_new;
      if (o != seen) continue;  // an id repeated inside the batch keeps the slot of its first occurrence
      ++seen;
      memcpy(s_ids + 16 * o, ids + 16 * i, 16);
      const Id128 key = load_id(ids + 16 * i);
      uint32_t mw = 0, ag = 0;
      auto om = h->orphan_meta.find(key);
      if (om != h->orphan_meta.end()) {
        mw = meta_word(true, om->second.first);
        ag = om->second.second;
      }
      s_meta[o] = mw;
      s_agent[o] = ag;
      s_seq[o] = seq ? seq[i] : 0;
    }
    CU(cudaMemcpyAsync(h->dIds + first_new * 16, s_ids, n_new * 16, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->dMeta + first_new, s_meta, n_new * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->dAgent + first_new, s_agent, n_new * 4, cudaMemcpyHostToDevice, s));
    if (h->with_seq) CU(cudaMemcpyAsync(h->dSeq + first_new, s_seq, n_new * 8, cudaMemcpyHostToDevice, s));
    launch_prepare_rows(h->dE, h->dNorm, h->dRnorm, h->dE16, h->dim, h->ld, h->ld16, (uint32_t)first_new,
                        (uint32_t)n_new, s);
    launch_count_irregular(h->dRnorm, h->dMeta, (uint32_t)first_new, (uint32_t)n_new, +1, h->d_irr, s);
    h->launches += h->dE16 ? 3 : 2;
  }
  for (uint32_t r : overwritten) {
    launch_prepare_rows(h->dE, h->dNorm, h->dRnorm, h->dE16, h->dim, h->ld, h->ld16, r, 1, s);
    launch_count_irregular(h->dRnorm, h->dMeta, r, 1, +1, h->d_irr, s);
    h->launches += h->dE16 ? 4 : 3;
  }
  CU(cudaGetLastError());

  // ---- 3. commit the host-side state
  if (n_new) {
    h->h_ids.insert(h->h_ids.end(), s_ids, s_ids + 16 * n_new);
    h->h_meta.insert(h->h_meta.end(), s_meta, s_meta + n_new);
    h->h_agent.insert(h->h_agent.end(), s_agent, s_agent + n_new);
    if (h->with_seq) h->h_seq.insert(h->h_seq.end(), s_seq, s_seq + n_new);
    for (uint64_t o = 0; o < n_new; ++o) {
      const Id128 key = load_id(s_ids + 16 * o);
      h->id2row.emplace(key, (uint32_t)(first_new + o));
      if (s_meta[o]) h->orphan_meta.erase(key);
    }
    h->n_rows += n_new;
    h->n_live += n_new;
    grow_ahead(h, n_new);
  }
  return mutation_done(h);
}

extern "C" cx_status cx_insert_batch(cx_index* h, const uint8_t* ids, const float* rows, uint64_t n,
                                     uint32_t len) {
  cx::CallerDevice keep_callers_device;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (h->shards) return shard_insert(h, ids, rows, n, len, false);
  return index_insert(h, ids, rows, n, len, false, nullptr);
}

// Bulk upsert of rows that already live in device memory (row-major [n][len] f32): the
// "mmap'd vectors bulk-uploaded once" path of the north star without a host round trip.
extern "C" cx_status cx_insert_batch_device(cx_index* h, const uint8_t* ids, const float* d_rows, uint64_t n,
                                            uint32_t len) {
  cx::CallerDevice keep_callers_device;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (h->shards) return shard_insert(h, ids, d_rows, n, len, true);
  return index_insert(h, ids, d_rows, n, len, true, nullptr);
}

extern "C" cx_status cx_insert(cx_index* h, const uint8_t id[16], const float* embedding, uint32_t len) {
  cx::CallerDevice keep_callers_device;
  return cx_insert_batch(h, id, embedding, 1, len);
}

cx_status cx::index_remove(cx_index* h, const uint8_t id[16]) {
  Id128 key = load_id(id);
  h->orphan_meta.erase(key);
  const uint32_t* found = h->id2row.find(key);
  if (!found) return CX_OK;  // index.rs:316-323: never an error
  CU(cudaSetDevice(h->device));
  cx_status st = settle(h);
  if (st != CX_OK) return st;
  uint32_t r = *found;
  h->id2row.erase(key);
  launch_count_irregular(h->dRnorm, h->dMeta, r, 1, -1, h->d_irr, h->mut_stream);  // while the row still counts as live
  h->h_meta[r] |= META_DEAD;
  h->n_live--;
  st = stage(h, 4);
  if (st != CX_OK) return st;
  *(uint32_t*)h->stage_h = h->h_meta[r];
  CU(cudaMemcpyAsync(h->dMeta + r, h->stage_h, 4, cudaMemcpyHostToDevice, h->mut_stream));
  h->launches += 1;
  return mutation_done(h);
}

extern "C" cx_status cx_remove(cx_index* h, const uint8_t id[16]) {
  cx::CallerDevice keep_callers_device;
  if (!h || !id) return fail(CX_ERR_VALIDATION, "null argument");
  if (h->shards) return shard_remove(h, id);
  return index_remove(h, id);
}

cx_status cx::index_set_metadata(cx_index* h, const uint8_t id[16], const char* kind, const char* agent) {
  uint32_t k = h->kinds.intern(kind);
  if (k > META_KIND_MASK) return fail(CX_ERR_VALIDATION, "more than 256 distinct node kinds");
  uint32_t a = h->agents.intern(agent);
  Id128 key = load_id(id);
  const uint32_t* found = h->id2row.find(key);
  if (!found) {  // the metadata map is independent of the vector map (index.rs:219-222)
    h->orphan_meta[key] = {k, a};
    return CX_OK;
  }
  CU(cudaSetDevice(h->device));
  cx_status st = settle(h);
  if (st != CX_OK) return st;
  uint32_t r = *found;
  h->h_meta[r] = meta_word(true, k) | (h->h_meta[r] & META_DEAD);
  h->h_agent[r] = a;
  st = stage(h, 8);
  if (st != CX_OK) return st;
  uint32_t* sp = (uint32_t*)h->stage_h;
  sp[0] = h->h_meta[r];
  sp[1] = a;
  CU(cudaMemcpyAsync(h->dMeta + r, sp, 4, cudaMemcpyHostToDevice, h->mut_stream));
  CU(cudaMemcpyAsync(h->dAgent + r, sp + 1, 4, cudaMemcpyHostToDevice, h->mut_stream));
  return mutation_done(h);
}

extern "C" cx_status cx_set_metadata(cx_index* h, const uint8_t id[16], const char* kind, const char* agent) {
  cx::CallerDevice keep_callers_device;
  if (!h || !id || !kind || !agent) return fail(CX_ERR_VALIDATION, "null argument");
  if (h->shards) return shard_set_metadata(h, id, kind, agent);
  return index_set_metadata(h, id, kind, agent);
}

extern "C" uint64_t cx_len(const cx_index* h) {
  cx::CallerDevice keep_callers_device;
  if (!h) return 0;
  return h->shards ? shard_len(h) : h->n_live;
}
extern "C" uint32_t cx_dimension(const cx_index* h) { return h ? h->dim : 0; }

// rebuild (index.rs:416-435): there is no graph to build; removed rows are compacted out of the store,
// order preserving and IN PLACE.  Live rows only ever move towards lower row numbers, so the store is
// processed in blocks of destination rows: a block's live sources are gathered into a bounce buffer and
// written back; every source a later block needs lies above everything written so far.
cx_status cx::index_rebuild(cx_index* h) {
  if (h->n_live == h->n_rows) return CX_OK;
  CU(cudaSetDevice(h->device));
  cx_status st = settle(h);
  if (st != CX_OK) return st;
  join_grower(h);
  cudaStream_t s = h->mut_stream;
  std::vector<uint32_t> live;
  live.reserve(h->n_live);
  for (uint32_t r = 0; r < h->n_rows; ++r)
    if (!(h->h_meta[r] & META_DEAD)) live.push_back(r);
  const uint64_t nl = live.size();
  uint64_t w_first = 0;  // rows below the first removed one stay where they are
  while (w_first < nl && live[w_first] == w_first) ++w_first;
  if (w_first < nl) {
    const uint64_t BLK = 32768;
    const uint64_t blk = nl - w_first < BLK ? nl - w_first : BLK;
    const size_t o_E = 0, o_nm = align_up(o_E + blk * h->ld * 4, 256), o_rn = align_up(o_nm + blk * 4, 256),
                 o_me = align_up(o_rn + blk * 4, 256), o_ag = align_up(o_me + blk * 4, 256),
                 o_id = align_up(o_ag + blk * 4, 256), o_16 = align_up(o_id + blk * 16, 256),
                 o_sq = align_up(o_16 + (h->dE16 ? blk * h->ld16 * 2 : 0), 256),
                 o_lv = align_up(o_sq + (h->dSeq ? blk * 8 : 0), 256), total = o_lv + (nl - w_first) * 4;
    char* b = nullptr;
    CU(cudaMalloc((void**)&b, total));
    std::unique_ptr<char, void (*)(char*)> guard(b, [](char* p) { cudaFree(p); });
    uint32_t* dlive = (uint32_t*)(b + o_lv);
    CU(cudaMemcpyAsync(dlive, live.data() + w_first, (nl - w_first) * 4, cudaMemcpyHostToDevice, s));
    const StoreView src = h->view();
    for (uint64_t w0 = w_first; w0 < nl; w0 += blk) {
      const uint64_t m = nl - w0 < blk ? nl - w0 : blk;
      launch_gather_rows(src, h->dSeq, (float*)(b + o_E), (float*)(b + o_nm), (float*)(b + o_rn), (uint32_t*)(b + o_me),
                         (uint32_t*)(b + o_ag), (uint8_t*)(b + o_id), h->dE16 ? (void*)(b + o_16) : nullptr,
                         h->dSeq ? (uint64_t*)(b + o_sq) : nullptr, dlive + (w0 - w_first), (uint32_t)m, s);
      CU(cudaMemcpyAsync(h->dE + w0 * h->ld, b + o_E, m * h->ld * 4, cudaMemcpyDeviceToDevice, s));
      CU(cudaMemcpyAsync(h->dNorm + w0, b + o_nm, m * 4, cudaMemcpyDeviceToDevice, s));
      CU(cudaMemcpyAsync(h->dRnorm + w0, b + o_rn, m * 4, cudaMemcpyDeviceToDevice, s));
      CU(cudaMemcpyAsync(h->dMeta + w0, b + o_me, m * 4, cudaMemcpyDeviceToDevice, s));
      CU(cudaMemcpyAsync(h->dAgent + w0, b + o_ag, m * 4, cudaMemcpyDeviceToDevice, s));
      CU(cudaMemcpyAsync(h->dIds + w0 * 16, b + o_id, m * 16, cudaMemcpyDeviceToDevice, s));
      if (h->dE16)
        CU(cudaMemcpyAsync((char*)h->dE16 + w0 * h->ld16 * 2, b + o_16, m * h->ld16 * 2, cudaMemcpyDeviceToDevice, s));
      if (h->dSeq) CU(cudaMemcpyAsync(h->dSeq + w0, b + o_sq, m * 8, cudaMemcpyDeviceToDevice, s));
      h->launches += 1;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
  }
  std::vector<uint8_t> nids(nl * 16);
  std::vector<uint32_t> nmeta(nl), nagent(nl);
  std::vector<uint64_t> nseq(h->with_seq ? nl : 0);
  h->id2row.clear();
  for (uint64_t i = 0; i < nl; ++i) {
    memcpy(&nids[i * 16], &h->h_ids[(size_t)live[i] * 16], 16);
    nmeta[i] = h->h_meta[live[i]];
    nagent[i] = h->h_agent[live[i]];
    if (h->with_seq) nseq[i] = h->h_seq[live[i]];
    h->id2row.emplace(load_id(&nids[i * 16]), (uint32_t)i);
  }
  h->h_ids.swap(nids);
  h->h_meta.swap(nmeta);
  h->h_agent.swap(nagent);
  h->h_seq.swap(nseq);
  h->n_rows = nl;
  h->n_live = nl;
  // give back the memory behind the removed rows (whole chunks above the live part)
  if (h->aE.vmm) {
    const uint64_t keep = nl > 1024 ? nl : 1024;
    h->aE.shrink(keep * h->ld * 4);
    h->aNorm.shrink((keep + 32) * 4);
    h->aRnorm.shrink((keep + 32) * 4);
    h->aMeta.shrink(keep * 4);
    h->aAgent.shrink(keep * 4);
    h->aIds.shrink(keep * 16);
    if (h->want_shadow) h->aE16.shrink(keep * h->ld16 * 2);
    if (h->with_seq) h->aSeq.shrink(keep * 8);
    // the usable capacity is what every array still has mapped
    uint64_t c = h->aE.mapped / ((size_t)h->ld * 4);
    auto lim = [&](uint64_t v) { if (v < c) c = v; };
    lim(h->aNorm.mapped / 4 - 32);
    lim(h->aRnorm.mapped / 4 - 32);
    lim(h->aMeta.mapped / 4);
    lim(h->aAgent.mapped / 4);
    lim(h->aIds.mapped / 16);
    if (h->want_shadow) lim(h->aE16.mapped / ((size_t)h->ld16 * 2));
    if (h->with_seq) lim(h->aSeq.mapped / 8);
    h->cap = c;
  }
  return mutation_done(h);
}

extern "C" cx_status cx_rebuild(cx_index* h) {
  cx::CallerDevice keep_callers_device;
  if (!h) return fail(CX_ERR_VALIDATION, "null index");
  if (h->shards) return shard_rebuild(h);
  return index_rebuild(h);
}

extern "C" cx_status cx_row_id(const cx_index* h, uint32_t row, uint8_t out_id[16]) {
  cx::CallerDevice keep_callers_device;
  if (!h || !out_id) return fail(CX_ERR_VALIDATION, "null argument");
  if (h->shards) return fail(CX_ERR_VALIDATION, "cx_row_id: a multi-device index returns ids, not rows");
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

// Writers shared with the multi-device index (which walks its shards in insertion order).
cx_status cx::index_save_vectors(cx_index* h, FILE* fp, uint64_t r0, uint64_t r1) {
  CU(cudaSetDevice(h->device));
  cx_status st = settle(h);
  if (st != CX_OK) return st;
  const uint64_t chunk = 65536;
  std::vector<float> buf((size_t)(r1 - r0 < chunk ? r1 - r0 : chunk) * h->ld);
  bool ok = true;
  for (uint64_t c0 = r0; c0 < r1 && ok; c0 += chunk) {
    uint64_t n = r1 - c0 < chunk ? r1 - c0 : chunk;
    if (cudaMemcpy(buf.data(), h->dE + c0 * h->ld, n * h->ld * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
      return fail(CX_ERR_CUDA, "device read failed during save");
    for (uint64_t i = 0; i < n && ok; ++i) {
      uint32_t r = (uint32_t)(c0 + i);
      if (h->h_meta[r] & META_DEAD) continue;
      ok = w64(fp, 16) && fwrite(&h->h_ids[(size_t)r * 16], 16, 1, fp) == 1 && w64(fp, h->dim) &&
           fwrite(&buf[i * h->ld], 4, h->dim, fp) == h->dim;
    }
  }
  return ok ? CX_OK : fail(CX_ERR_IO, "Failed to write index file");
}

uint64_t cx::index_meta_count(const cx_index* h) {
  uint64_t n_meta = h->orphan_meta.size();
  for (uint32_t r = 0; r < h->n_rows; ++r)
    if (!(h->h_meta[r] & META_DEAD) && (h->h_meta[r] & META_HAS)) ++n_meta;
  return n_meta;
}

bool cx::index_save_meta(const cx_index* h, FILE* fp) {
  bool ok = true;
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
  return ok;
}

extern "C" cx_status cx_save(const cx_index* hc, const char* path) {
  cx::CallerDevice keep_callers_device;
  cx_index* h = const_cast<cx_index*>(hc);
  if (!h || !path) return fail(CX_ERR_VALIDATION, "null argument");
  if (h->shards) return shard_save(h, path);
  FILE* fp = fopen(path, "wb");
  if (!fp) return fail(CX_ERR_IO, "Failed to write index file: %s", path);
  bool ok = w64(fp, h->n_live);
  if (ok && index_save_vectors(h, fp, 0, h->n_rows) != CX_OK) ok = false;
  ok = ok && w64(fp, index_meta_count(h)) && index_save_meta(h, fp);
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

// Reads the file into an index created by `make(dim, &h)` (one device or several).
cx_status cx::index_load_into(const char* path, cx_status (*make)(uint32_t, void*, cx_index**), void* ctx,
                              cx_index** out) {
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
  cx_status st = make((uint32_t)dim, ctx, &h);
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

extern "C" cx_status cx_load(const char* path, int device, cx_index** out) {
  cx::CallerDevice keep_callers_device;
  if (!path || !out) return fail(CX_ERR_VALIDATION, "null argument");
  auto make = [](uint32_t dim, void* ctx, cx_index** o) { return cx_index_create(dim, *(int*)ctx, o); };
  return index_load_into(path, make, &device, out);
}

// ------------------------------------------------------------------------------
void cx::index_add_stats(const cx_index* h, cx_stats* out) {
  out->kernel_launches += h->launches.load();
  out->queries_stream += h->q_stream.load();
  out->queries_stream_bf16 += h->q_stream16.load();
  out->queries_tensor += h->q_tensor.load();
  out->queries_exact += h->q_exact.load();
  out->fallbacks += h->fallbacks.load();
  out->h2d_bytes += h->h2d.load();
  out->d2h_bytes += h->d2h.load();
  out->pass_kernel_ns += h->pass_ns.load();
  out->pass_kernel_launches += h->pass_launches.load();
  out->graph_launches += h->graph_launches.load();
  out->grow_events += h->grow_events.load();
  out->irregular_rows += h->n_irregular();
  out->capacity_rows += h->cap.load();
  out->grow_ns += h->grow_ns.load();
  if (h->grow_ns_max.load() > out->grow_ns_max) out->grow_ns_max = h->grow_ns_max.load();
  out->grow_waits += h->grow_waits.load();
  out->in_place_growth = h->aE.vmm ? 1 : 0;
}

extern "C" cx_status cx_get_stats(const cx_index* hc, cx_stats* out) {
  cx::CallerDevice keep_callers_device;
  cx_index* h = const_cast<cx_index*>(hc);
  if (!h || !out) return fail(CX_ERR_VALIDATION, "null argument");
  memset(out, 0, sizeof *out);
  if (h->shards) return shard_stats(h, out);
  cx_status st = settle(h);  // the irregular-row count of the last mutation
  if (st != CX_OK) return st;
  index_add_stats(h, out);
  CU(cudaSetDevice(h->device));
  uint64_t why[3];
  select_why_read(why);
  out->unverified_overflow = why[0];
  out->unverified_near_ties = why[1];
  out->unverified_other = why[2];
  return CX_OK;
}

cx_status cx::index_set_option(cx_index* h, const char* key, int64_t value) {
  if (!strcmp(key, "force_path")) {
    if (value < 0 || value > 3) return fail(CX_ERR_VALIDATION, "force_path must be 0..3");
    h->force_path = (int)value;
    return CX_OK;
  }
  if (!strcmp(key, "profile")) {
    h->profile = value != 0;
    return CX_OK;
  }
  if (!strcmp(key, "graphs")) {
    h->use_graphs = value != 0;
    return CX_OK;
  }
  if (!strcmp(key, "tensor_pair")) {  // CTA-pair (cta_group::2) form of the tensor pass for batches of two or more query tiles
    h->tensor_tune.pair = value != 0;
    return CX_OK;
  }
  if (!strcmp(key, "stream_bf16")) {  // 0: small batches stream the fp32 rows even when a bf16 shadow exists
    h->stream_bf16 = value != 0;
    return CX_OK;
  }
  if (!strcmp(key, "tensor_leftover_sms")) {  // 0: only the regular (query tiles x row splits) grid of CTAs
    h->tensor_tune.use_leftover_sms = value != 0;
    return CX_OK;
  }
  if (!strcmp(key, "tensor_phase_growth")) {
    if (value < 0 || value > 1024) return fail(CX_ERR_VALIDATION, "tensor_phase_growth must be 0..1024");
    h->tensor_phase_growth = (uint32_t)value;
    return CX_OK;
  }
  if (!strcmp(key, "tensor_sample_tiles")) {
    if (value < 0 || value > 4096) return fail(CX_ERR_VALIDATION, "tensor_sample_tiles must be 0..4096");
    h->tensor_sample_tiles = (uint32_t)value;
    return CX_OK;
  }
  if (!strcmp(key, "blocking_sync")) {
    h->blocking_sync = value != 0;
    return CX_OK;
  }
  if (!strcmp(key, "tensor_epi_warps")) {
    if (value != 8 && value != 16) return fail(CX_ERR_VALIDATION, "tensor_epi_warps must be 8 or 16");
    h->tensor_tune.epi_warps = (int)value;
    return CX_OK;
  }
#ifdef CX_PROBE
  if (!strcmp(key, "tensor_debug")) {  // measurement hook of probe builds only: results of the tensor pass become wrong
    if (value == -1) tensor_report_clock();
    else h->tensor_tune.debug = (int)value;
    return CX_OK;
  }
#endif
  if (!strcmp(key, "tensor_min_batch")) {
    if (value < 1) return fail(CX_ERR_VALIDATION, "tensor_min_batch must be >= 1");
    h->tensor_min_batch = (uint32_t)value;
    return CX_OK;
  }
  if (!strcmp(key, "shadow")) {  // bf16 shadow matrix for the tensor pass; only before the first insert
    if (h->store_ready) return fail(CX_ERR_VALIDATION, "shadow can only be changed on an empty index");
    h->want_shadow = value != 0;
    return CX_OK;
  }
  return fail(CX_ERR_VALIDATION, "unknown option %s", key);
}

extern "C" cx_status cx_set_option(cx_index* h, const char* key, int64_t value) {
  cx::CallerDevice keep_callers_device;
  if (!h || !key) return fail(CX_ERR_VALIDATION, "null argument");
  if (h->shards) return shard_set_option(h, key, value);
  return index_set_option(h, key, value);
}

extern "C" const char* cx_last_error(void) { return cx::last_error(); }
extern "C" const char* cx_version(void) { return "cortex_b200 0.2.0 (sm_100a)"; }
