// The reference's own unit tests for the vector index
// (/root/reference/crates/cortex-core/src/vector/index.rs:475-729), in C++, against
// cortex::GpuVectorIndex (cortex_b200/host/vector_index.hpp) over the C ABI.
// Built and run by tests/test_gpu_cpp_host.py on a GPU box.
#include <cassert>
#include <cmath>
#include <cstdio>
#include <random>

#include "../../cortex_b200/host/vector_index.hpp"

using namespace cortex;

static NodeId new_id() {
  static std::mt19937_64 rng(0xC027E5);
  NodeId id;
  uint64_t a = rng(), b = rng();
  __builtin_memcpy(id.data(), &a, 8);
  __builtin_memcpy(id.data() + 8, &b, 8);
  return id;
}
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int test_index_insert_and_search() {  // index.rs:484-510
  GpuVectorIndex ix(3);
  NodeId id1 = new_id(), id2 = new_id(), id3 = new_id();
  ix.insert(id1, {1.0f, 0.0f, 0.0f});
  ix.insert(id2, {0.9f, 0.1f, 0.0f});
  ix.insert(id3, {0.0f, 1.0f, 0.0f});
  ix.rebuild();
  auto r = ix.search({1.0f, 0.0f, 0.0f}, 2);
  CHECK(r.size() == 2);
  CHECK(r[0].node_id == id1);
  return 0;
}
static int test_threshold_search() {  // :513-535
  GpuVectorIndex ix(3);
  NodeId id1 = new_id(), id2 = new_id();
  ix.insert(id1, {1.0f, 0.0f, 0.0f});
  ix.insert(id2, {0.0f, 1.0f, 0.0f});
  ix.rebuild();
  auto r = ix.search_threshold({1.0f, 0.0f, 0.0f}, 0.95f);
  CHECK(r.size() == 1 && r[0].node_id == id1);
  return 0;
}
static int test_index_persistence(const char* dir) {  // :538-566
  std::string path = std::string(dir) + "/test.hnsw";
  GpuVectorIndex ix(3);
  NodeId id1 = new_id();
  ix.insert(id1, {1.0f, 0.0f, 0.0f});
  ix.rebuild();
  ix.save(path);
  GpuVectorIndex ld = GpuVectorIndex::load(path);
  CHECK(ld.len() == 1);
  auto r = ld.search({1.0f, 0.0f, 0.0f}, 1);
  CHECK(r.size() == 1 && r[0].node_id == id1);
  return 0;
}
static int test_dimension_mismatch_rejected() {  // :579-583
  GpuVectorIndex ix(3);
  bool threw = false;
  try { ix.insert(new_id(), {1.0f, 2.0f}); } catch (const CortexError& e) {
    threw = std::string(e.what()).find("Embedding dimension mismatch: expected 3, got 2") != std::string::npos;
  }
  CHECK(threw);
  return 0;
}
static int test_empty_index_search() {  // :586-590
  GpuVectorIndex ix(3);
  CHECK(ix.search({1.0f, 0.0f, 0.0f}, 5).empty());
  CHECK(ix.is_empty());
  return 0;
}
static int test_brute_force_fallback() {  // :593-606
  GpuVectorIndex ix(3);
  NodeId id1 = new_id(), id2 = new_id();
  ix.insert(id1, {1.0f, 0.0f, 0.0f});
  ix.insert(id2, {0.0f, 1.0f, 0.0f});
  auto r = ix.search({1.0f, 0.0f, 0.0f}, 2);
  CHECK(r.size() == 2 && r[0].node_id == id1);
  return 0;
}
static int test_filter_by_kind() {  // :609-627
  GpuVectorIndex ix = GpuVectorIndex::with_metadata(3);
  NodeId id1 = new_id(), id2 = new_id();
  ix.insert(id1, {1.0f, 0.0f, 0.0f});
  ix.set_metadata(id1, "fact", "test");
  ix.insert(id2, {0.9f, 0.1f, 0.0f});
  ix.set_metadata(id2, "decision", "test");
  ix.rebuild();
  VectorFilter f;
  f.with_kinds({"decision"});
  auto r = ix.search({1.0f, 0.0f, 0.0f}, 5, &f);
  CHECK(r.size() == 1 && r[0].node_id == id2);
  return 0;
}
static int test_filter_exclude() {  // :630-646
  GpuVectorIndex ix(3);
  NodeId id1 = new_id(), id2 = new_id();
  ix.insert(id1, {1.0f, 0.0f, 0.0f});
  ix.insert(id2, {0.9f, 0.1f, 0.0f});
  ix.rebuild();
  VectorFilter f;
  f.excluding({id1});
  auto r = ix.search({1.0f, 0.0f, 0.0f}, 5, &f);
  CHECK(r.size() == 1 && r[0].node_id == id2);
  return 0;
}
static int test_remove_doesnt_crash_search() {  // :649-664
  GpuVectorIndex ix(3);
  NodeId id1 = new_id(), id2 = new_id();
  ix.insert(id1, {1.0f, 0.0f, 0.0f});
  ix.insert(id2, {0.0f, 1.0f, 0.0f});
  ix.rebuild();
  ix.remove(id1);
  CHECK(ix.len() == 1);
  CHECK(!ix.search({1.0f, 0.0f, 0.0f}, 5).empty());
  return 0;
}
static int test_search_batch() {  // :667-684
  GpuVectorIndex ix(3);
  NodeId id1 = new_id(), id2 = new_id(), id3 = new_id();
  ix.insert(id1, {1.0f, 0.0f, 0.0f});
  ix.insert(id2, {0.0f, 1.0f, 0.0f});
  ix.insert(id3, {0.0f, 0.0f, 1.0f});
  ix.rebuild();
  auto res = ix.search_batch({{id1, {1.0f, 0.0f, 0.0f}}, {id2, {0.0f, 1.0f, 0.0f}}}, 1);
  CHECK(res.size() == 2);
  CHECK(res[id1][0].node_id == id1 && res[id2][0].node_id == id2);
  return 0;
}
static int test_similarity_score_range() {  // :687-708
  GpuVectorIndex ix(3);
  ix.insert(new_id(), {1.0f, 0.0f, 0.0f});
  ix.insert(new_id(), {-1.0f, 0.0f, 0.0f});
  ix.rebuild();
  auto r = ix.search({1.0f, 0.0f, 0.0f}, 2);
  for (auto& x : r) CHECK(x.score >= 0.0f && x.score <= 1.0f);
  CHECK(r[0].score > 0.99f);
  return 0;
}
static int test_threshold_returns_only_above() {  // :711-728
  GpuVectorIndex ix(3);
  NodeId idc = new_id(), idf = new_id();
  ix.insert(idc, {1.0f, 0.0f, 0.0f});
  ix.insert(idf, {0.0f, 0.0f, 1.0f});
  ix.rebuild();
  auto r = ix.search_threshold({1.0f, 0.0f, 0.0f}, 0.5f);
  bool any = false;
  for (auto& x : r) { CHECK(x.score >= 0.5f); any = any || x.node_id == idc; }
  CHECK(any);
  return 0;
}

int main(int argc, char** argv) {
  const char* dir = argc > 1 ? argv[1] : "/tmp";
  int fails = 0;
  fails += test_index_insert_and_search();
  fails += test_threshold_search();
  fails += test_index_persistence(dir);
  fails += test_dimension_mismatch_rejected();
  fails += test_empty_index_search();
  fails += test_brute_force_fallback();
  fails += test_filter_by_kind();
  fails += test_filter_exclude();
  fails += test_remove_doesnt_crash_search();
  fails += test_search_batch();
  fails += test_similarity_score_range();
  fails += test_threshold_returns_only_above();
  std::printf("%s (%d failed of 12)\n", fails ? "FAILED" : "ok", fails);
  return fails ? 1 : 0;
}
