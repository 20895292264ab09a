// vector_index.hpp -- C++ host-side mirror of cortex-core's vector-index surface over the
// C ABI (include/cortex_gpu.h).  The reference is compiled code (Rust); with no Rust
// toolchain in the build image this header plays the role of the reference-language host
// layer: same names, argument meaning and error behaviour as
//   /root/reference/crates/cortex-core/src/vector/index.rs
//     SimilarityResult :9-15   VectorFilter :17-47   trait VectorIndex :50-99
//     HnswIndex::new :204-211  set_metadata :219-222
// Header only; link against libcortex_gpu.so.  There is no fallback implementation.
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/cortex_gpu.h"

namespace cortex {

using NodeId = std::array<uint8_t, 16>;   // uuid::Uuid bytes (types.rs:9)
using Embedding = std::vector<float>;     // types.rs:22

struct NodeIdHash {
  size_t operator()(const NodeId& id) const {
    uint64_t a, b;
    __builtin_memcpy(&a, id.data(), 8);
    __builtin_memcpy(&b, id.data() + 8, 8);
    return (size_t)(a * 0x9E3779B97F4A7C15ull ^ b);
  }
};

// CortexError::Validation(String) (error.rs:48-49) -- the only variant the index raises
struct CortexError : std::runtime_error {
  int status;
  CortexError(const std::string& msg, int st) : std::runtime_error("Validation error: " + msg), status(st) {}
};

struct SimilarityResult {
  NodeId node_id;
  float score;     // cosine similarity clamped to [0,1]
  float distance;  // 1 - similarity (unclamped)
};

struct VectorFilter {
  std::optional<std::vector<std::string>> kinds;
  std::optional<std::vector<NodeId>> exclude;
  std::optional<std::string> source_agent;
  VectorFilter& with_kinds(std::vector<std::string> k) { kinds = std::move(k); return *this; }
  VectorFilter& excluding(std::vector<NodeId> ids) { exclude = std::move(ids); return *this; }
  VectorFilter& with_source_agent(std::string a) { source_agent = std::move(a); return *this; }
};

// trait VectorIndex (index.rs:50-99)
class VectorIndex {
 public:
  virtual ~VectorIndex() = default;
  virtual void insert(const NodeId& id, const Embedding& embedding) = 0;
  virtual void remove(const NodeId& id) = 0;
  virtual std::vector<SimilarityResult> search(const Embedding& query, size_t k,
                                               const VectorFilter* filter = nullptr) const = 0;
  virtual std::vector<SimilarityResult> search_threshold(const Embedding& query, float threshold,
                                                         const VectorFilter* filter = nullptr) const = 0;
  virtual std::unordered_map<NodeId, std::vector<SimilarityResult>, NodeIdHash> search_batch(
      const std::vector<std::pair<NodeId, Embedding>>& queries, size_t k,
      const VectorFilter* filter = nullptr) const = 0;
  virtual size_t len() const = 0;
  bool is_empty() const { return len() == 0; }
  virtual void rebuild() = 0;
  virtual void save(const std::string& path) const = 0;
};

class GpuVectorIndex final : public VectorIndex {
 public:
  explicit GpuVectorIndex(size_t dimension, int device = 0) {
    check(cx_index_create((uint32_t)dimension, device, &h_));
  }
  // one index row-sharded over several GPUs of the box, driven from this process
  GpuVectorIndex(size_t dimension, const std::vector<int>& devices) {
    check(cx_index_create_sharded((uint32_t)dimension, devices.data(), (uint32_t)devices.size(), &h_));
  }
  static GpuVectorIndex with_metadata(size_t dimension) { return GpuVectorIndex(dimension); }
  static GpuVectorIndex load(const std::string& path, const std::vector<int>& devices) {
    cx_index* h = nullptr;
    check(cx_load_sharded(path.c_str(), devices.data(), (uint32_t)devices.size(), &h));
    return GpuVectorIndex(h);
  }
  size_t shard_count() const { return cx_shard_count(h_); }
  static GpuVectorIndex load(const std::string& path, int device = 0) {
    cx_index* h = nullptr;
    check(cx_load(path.c_str(), device, &h));
    return GpuVectorIndex(h);
  }
  GpuVectorIndex(GpuVectorIndex&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  GpuVectorIndex(const GpuVectorIndex&) = delete;
  ~GpuVectorIndex() override { if (h_) cx_index_destroy(h_); }

  void set_metadata(const NodeId& id, const std::string& kind, const std::string& source_agent) {
    check(cx_set_metadata(h_, id.data(), kind.c_str(), source_agent.c_str()));
  }
  void insert(const NodeId& id, const Embedding& e) override {
    check(cx_insert(h_, id.data(), e.data(), (uint32_t)e.size()));
  }
  void remove(const NodeId& id) override { check(cx_remove(h_, id.data())); }
  size_t len() const override { return (size_t)cx_len(h_); }
  void rebuild() override { check(cx_rebuild(h_)); }
  void save(const std::string& path) const override { check(cx_save(h_, path.c_str())); }

  std::vector<SimilarityResult> search(const Embedding& q, size_t k, const VectorFilter* f = nullptr) const override {
    const size_t kk = std::max<size_t>(1, std::min(k, std::max<size_t>(1, len())));
    std::vector<uint8_t> ids(16 * kk);
    std::vector<float> sc(kk), di(kk);
    uint64_t n = 0;
    CFilter cf(f);
    check(cx_search(h_, q.data(), (uint32_t)q.size(), std::min(k, kk), cf.ptr(), ids.data(), sc.data(), di.data(), &n));
    return collect(ids.data(), sc.data(), di.data(), n);
  }
  std::vector<SimilarityResult> search_threshold(const Embedding& q, float threshold,
                                                 const VectorFilter* f = nullptr) const override {
    CFilter cf(f);
    size_t cap = std::min<size_t>(std::max<size_t>(1, len()), 4096);
    for (;;) {
      std::vector<uint8_t> ids(16 * cap);
      std::vector<float> sc(cap), di(cap);
      uint64_t n = 0, total = 0;
      check(cx_search_threshold(h_, q.data(), (uint32_t)q.size(), threshold, cf.ptr(), cap, ids.data(), sc.data(),
                                di.data(), &n, &total));
      if (total <= cap) return collect(ids.data(), sc.data(), di.data(), n);
      cap = (size_t)total;
    }
  }
  std::unordered_map<NodeId, std::vector<SimilarityResult>, NodeIdHash> search_batch(
      const std::vector<std::pair<NodeId, Embedding>>& queries, size_t k,
      const VectorFilter* f = nullptr) const override {
    std::unordered_map<NodeId, std::vector<SimilarityResult>, NodeIdHash> out;
    if (queries.empty()) return out;
    // the reference searches every query on its own (index.rs:397-403), so lengths may differ: one call
    // per distinct length (normally exactly one), each over a flat buffer of that stride
    k = std::max<size_t>(1, std::min(k, std::max<size_t>(1, len())));
    std::unordered_map<size_t, std::vector<size_t>> by_len;
    for (size_t i = 0; i < queries.size(); ++i) by_len[queries[i].second.size()].push_back(i);
    CFilter cf(f);
    for (auto& g : by_len) {
      const size_t dim = g.first, B = g.second.size();
      std::vector<float> flat(B * dim);
      for (size_t b = 0; b < B; ++b) {
        const Embedding& e = queries[g.second[b]].second;
        std::copy(e.begin(), e.end(), flat.begin() + b * dim);
      }
      std::vector<uint8_t> ids(16 * B * k);
      std::vector<float> sc(B * k), di(B * k);
      std::vector<uint64_t> n(B);
      check(cx_search_batch(h_, flat.data(), B, (uint32_t)dim, k, cf.ptr(), ids.data(), sc.data(), di.data(), n.data()));
      for (size_t b = 0; b < B; ++b)
        out[queries[g.second[b]].first] = collect(ids.data() + 16 * b * k, sc.data() + b * k, di.data() + b * k, n[b]);
    }
    return out;
  }
  // The similarity step of DedupScanner::scan (linker/dedup.rs:65-127) as one self-join: every unordered
  // pair of live nodes with score >= threshold, once, a inserted before b.
  struct DuplicatePair {
    NodeId node_a, node_b;
    float similarity;
  };
  std::vector<DuplicatePair> dedup_scan(float threshold, uint32_t per_node_cap = 64, size_t max_pairs = 1 << 20,
                                        uint64_t* total = nullptr) const {
    std::vector<uint8_t> a(16 * max_pairs), b(16 * max_pairs);
    std::vector<float> sc(max_pairs);
    uint64_t n = 0, tot = 0;
    check(cx_dedup_scan(h_, threshold, per_node_cap, max_pairs, a.data(), b.data(), sc.data(), &n, &tot));
    if (total) *total = tot;
    std::vector<DuplicatePair> out(n);
    for (uint64_t i = 0; i < n; ++i) {
      __builtin_memcpy(out[i].node_a.data(), a.data() + 16 * i, 16);
      __builtin_memcpy(out[i].node_b.data(), b.data() + 16 * i, 16);
      out[i].similarity = sc[i];
    }
    return out;
  }
  cx_index* handle() const { return h_; }

 private:
  explicit GpuVectorIndex(cx_index* h) : h_(h) {}
  cx_index* h_ = nullptr;

  static void check(int st) {
    if (st != CX_OK) throw CortexError(cx_last_error(), st);
  }
  static std::vector<SimilarityResult> collect(const uint8_t* ids, const float* sc, const float* di, uint64_t n) {
    std::vector<SimilarityResult> r(n);
    for (uint64_t i = 0; i < n; ++i) {
      __builtin_memcpy(r[i].node_id.data(), ids + 16 * i, 16);
      r[i].score = sc[i];
      r[i].distance = di[i];
    }
    return r;
  }
  struct CFilter {
    cx_filter raw{};
    bool present = false;
    std::vector<const char*> kind_ptrs;
    std::vector<uint8_t> excl;
    explicit CFilter(const VectorFilter* f) {
      if (!f) return;
      present = true;
      if (f->kinds) {
        for (auto& k : *f->kinds) kind_ptrs.push_back(k.c_str());
        raw.has_kinds = 1;
        raw.kinds = kind_ptrs.data();
        raw.n_kinds = (uint32_t)kind_ptrs.size();
      }
      if (f->exclude) {
        for (auto& id : *f->exclude) excl.insert(excl.end(), id.begin(), id.end());
        raw.has_exclude = 1;
        raw.exclude_ids = excl.data();
        raw.n_exclude = (uint32_t)f->exclude->size();
      }
      if (f->source_agent) {
        raw.has_source_agent = 1;
        raw.source_agent = f->source_agent->c_str();
      }
    }
    const cx_filter* ptr() const { return present ? &raw : nullptr; }
  };
};

}  // namespace cortex
