/*
 * cortex_gpu.h -- C ABI of the B200-native similarity-scan engine.
 *
 * This is the drop-in boundary for cortex-core's vector layer: one exported
 * function per method of `trait VectorIndex`
 * (/root/reference/crates/cortex-core/src/vector/index.rs:50-99) plus
 * `HnswIndex::new` / `set_metadata` (index.rs:204-222).  A Rust FFI crate
 * (rust/cortex-gpu-sys) binds exactly these symbols and
 * implements `VectorIndex for GpuVectorIndex` on top of them.
 *
 * Conventions
 *  - plain pointers and sizes only; ids are 16 raw bytes (uuid::Uuid::as_bytes,
 *    types.rs:9); embeddings are contiguous f32 (types.rs:22).
 *  - every function returns a cx_status; on failure cx_last_error() returns a
 *    thread-local message.  The Rust side maps any non-zero status to
 *    CortexError::Validation(msg) (error.rs:48-49), the only error the reference
 *    index ever raises (index.rs:299-305, 438-459).
 *  - search* functions are re-entrant from many host threads (the reference
 *    shares the index as Arc<RwLock<_>> and searches under read guards,
 *    serve.rs:101, routes.rs:906); mutators (insert/remove/set_metadata/rebuild)
 *    must be externally serialised against everything else, which the caller's
 *    write guard already does.
 *  - every function leaves the calling thread's current CUDA device as it found
 *    it (an index may live on any device, a multi-device handle on several); the
 *    stream / pointer arguments of the *_device functions belong to the index's
 *    device (devices[0] of a multi-device handle).
 *  - there is no CPU fallback: without a CUDA device cx_index_create fails with
 *    CX_ERR_CUDA.
 */
#ifndef CORTEX_GPU_H
#define CORTEX_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cx_status {
  CX_OK = 0,
  CX_ERR_VALIDATION = 1, /* dimension mismatch, bad argument (index.rs:299-305) */
  CX_ERR_CUDA = 2,
  CX_ERR_NCCL = 3,
  CX_ERR_IO = 4          /* save/load (index.rs:438-459) */
} cx_status;

typedef struct cx_index cx_index;

/* VectorFilter, index.rs:17-26.  A NULL filter pointer == None. */
typedef struct cx_filter {
  int32_t has_kinds;            /* kinds: Option<Vec<NodeKind>> */
  const char* const* kinds;
  uint32_t n_kinds;
  int32_t has_exclude;          /* exclude: Option<Vec<NodeId>> */
  const uint8_t* exclude_ids;   /* n_exclude x 16 bytes */
  uint32_t n_exclude;
  int32_t has_source_agent;     /* source_agent: Option<String> */
  const char* source_agent;
} cx_filter;

/* Counters for tests / bench (gpu_launches, which pass served a query). */
typedef struct cx_stats {
  uint64_t kernel_launches;     /* kernels launched by this index since creation */
  uint64_t queries_stream;      /* queries answered by the streaming pass (K1), fp32 rows or bf16 shadow */
  uint64_t queries_tensor;      /* queries answered by the tcgen05 pass (K2) */
  uint64_t queries_exact;       /* queries answered by the exact path */
  uint64_t fallbacks;           /* queries whose fast-pass result failed verification */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint64_t pass_kernel_ns;      /* with option "profile"=1: CUDA-event time of the scan-pass kernels */
  uint64_t pass_kernel_launches;/* ... and how many of them that time covers */
  uint64_t graph_launches;      /* device-resident searches replayed as one CUDA graph launch */
  uint64_t grow_events;         /* times the store was extended (in place: no copy) */
  uint64_t irregular_rows;      /* live rows whose norm under- / overflows fp32: while > 0 every search takes the exact path */
  uint64_t capacity_rows;       /* rows the mapped store can hold before it is extended again */
  uint64_t in_place_growth;     /* 1 = the store grows in place (driver virtual-memory API), 0 = allocate-copy-free */
  uint64_t grow_ns;             /* host time spent extending the store (mostly on the helper thread, ahead of need) */
  uint64_t grow_ns_max;         /* ... the longest single extension */
  uint64_t grow_waits;          /* inserts that had to wait for an extension */
  /* why fast-pass results failed verification (and were redone on a tighter path), on this device since
   * process start: candidate list overflow / near-ties around rank k denser than the rescored band / other */
  uint64_t unverified_overflow, unverified_near_ties, unverified_other;
  uint64_t queries_stream_bf16; /* of queries_stream: answered by the streaming pass over the bf16 shadow (B <= 4) */
} cx_stats;

/* HnswIndex::new(dimension), index.rs:204-211.  device = CUDA ordinal. */
cx_status cx_index_create(uint32_t dimension, int device, cx_index** out);
/* The same index row-sharded over n_devices GPUs of one box, driven from this one process.
 * The handle is used with every function below exactly like a single-device one: inserts are
 * distributed so that the shards stay balanced, search / search_threshold / search_batch /
 * autolink_batch / dedup_scan fan out to every device on its own streams and the per-device lists
 * are merged on devices[0] (written there over NVLink by the producing GPUs).  Results, including the order
 * of equal scores (global insertion order), are identical to a single-device index holding the same
 * rows.  A device may be listed more than once (several shards on one GPU: how the single-GPU tests
 * exercise this path). */
cx_status cx_index_create_sharded(uint32_t dimension, const int* devices, uint32_t n_devices, cx_index** out);
/* number of shards behind a handle (1 for cx_index_create) */
uint32_t cx_shard_count(const cx_index* h);
void cx_index_destroy(cx_index* h);

/* VectorIndex::insert, index.rs:298-314.  len != dimension -> CX_ERR_VALIDATION
 * "Embedding dimension mismatch: expected D, got L".  Same id overwrites. */
cx_status cx_insert(cx_index* h, const uint8_t id[16], const float* embedding, uint32_t len);
/* Bulk form of the startup loop serve.rs:111-117 / api.rs:55-69: n rows, row-major. */
cx_status cx_insert_batch(cx_index* h, const uint8_t* ids, const float* rows, uint64_t n, uint32_t len);
/* Same, for rows already resident in device memory ([n][len] f32, row-major; for a multi-device index:
 * in the memory of devices[0]).  ids stay host-side.  Synchronises the device first (the buffer may
 * have been produced on any stream). */
cx_status cx_insert_batch_device(cx_index* h, const uint8_t* ids, const float* d_rows, uint64_t n, uint32_t len);
/* VectorIndex::remove, index.rs:316-323.  Unknown id is not an error. */
cx_status cx_remove(cx_index* h, const uint8_t id[16]);
/* HnswIndex::set_metadata, index.rs:219-222. */
cx_status cx_set_metadata(cx_index* h, const uint8_t id[16], const char* kind, const char* source_agent);
/* VectorIndex::len, index.rs:412-414. */
uint64_t cx_len(const cx_index* h);
uint32_t cx_dimension(const cx_index* h);
/* VectorIndex::rebuild, index.rs:416-435: there is no graph to build; this
 * compacts removed rows out of the device matrix (order preserving, in place). */
cx_status cx_rebuild(cx_index* h);
/* Map capacity for n rows up front (no reference counterpart; optional: the store grows in place). */
cx_status cx_reserve(cx_index* h, uint64_t n_rows);

/* insert / remove / set_metadata / rebuild enqueue their device work and return; the next search (or
 * mutation, save, stats) on the handle waits for it.  All of them keep the host-side state unchanged
 * when they fail. */

/* VectorIndex::search, index.rs:325-374 (exact scan semantics of :259-294).
 * Outputs hold up to k entries; *out_n = number written.  qlen may differ from
 * the dimension (the reference never checks it; zip truncates, index.rs:172). */
cx_status cx_search(cx_index* h, const float* query, uint32_t qlen, uint64_t k, const cx_filter* filter,
                    uint8_t* out_ids, float* out_score, float* out_distance, uint64_t* out_n);
/* VectorIndex::search_threshold, index.rs:376-388: all rows with score >= threshold,
 * best first.  At most cap entries are written; *out_total = how many qualify. */
cx_status cx_search_threshold(cx_index* h, const float* query, uint32_t qlen, float threshold,
                              const cx_filter* filter, uint64_t cap, uint8_t* out_ids, float* out_score,
                              float* out_distance, uint64_t* out_n, uint64_t* out_total);
/* search_threshold for B queries at once (the call the dedup scanner and the link rules make
 * once per node, linker/dedup.rs:84-86): outputs are [B][cap], out_n[b] entries written,
 * out_total[b] rows qualify. */
cx_status cx_search_threshold_batch(cx_index* h, const float* queries, uint64_t B, uint32_t qlen, float threshold,
                                    const cx_filter* filter, uint64_t cap, uint8_t* out_ids, float* out_score,
                                    float* out_distance, uint64_t* out_n, uint64_t* out_total);
/* DedupScanner::scan, linker/dedup.rs:65-127, as one self-join: every live node runs
 * search_threshold(own embedding, threshold), skips itself and reports each unordered pair
 * once.  At most per_node_cap partners per node and max_pairs pairs are written (a-id, b-id,
 * score), ordered by (insertion order of a, score desc, insertion order of b) with a inserted
 * before b; *out_total = pairs that qualify.  Choosing merge / supersede / link for a pair
 * (determine_action, :130-171) needs graph degrees and stays with the caller.
 * Like the reference it always succeeds: the tcgen05 self-join serves thresholds >= 0.2 on indexes of
 * 256+ regular rows; a node with more than 2 048 near-threshold partners (a large duplicate cluster),
 * lower thresholds, tiny indexes and indexes holding rows with under- / overflowing norms are served
 * per node by the exact path (reference arithmetic over every row + sort: correct, but milliseconds
 * per node at a million rows).  Pairs are compacted on the device: the copy back is 12 bytes per pair. */
cx_status cx_dedup_scan(cx_index* h, float threshold, uint32_t per_node_cap, uint64_t max_pairs,
                        uint8_t* out_a_ids, uint8_t* out_b_ids, float* out_score, uint64_t* out_n,
                        uint64_t* out_total);
/* VectorIndex::search_batch, index.rs:390-410.  queries is [B][qlen] row-major;
 * outputs are [B][k] (ids [B][k][16]); out_n[b] = results of query b.  The caller
 * keys the result map by its own query ids. */
cx_status cx_search_batch(cx_index* h, const float* queries, uint64_t B, uint32_t qlen, uint64_t k,
                          const cx_filter* filter, uint8_t* out_ids, float* out_score, float* out_distance,
                          uint64_t* out_n);

/* Device-resident form of search_batch for callers that keep queries and results
 * in HBM (sharded merge over NCCL, benchmarks).  d_queries: [B][qlen] f32 device
 * memory; outputs are device buffers [B][k] (rows are shard-local row numbers),
 * d_out_n [B] u32.  Runs on `stream` (a cudaStream_t) and returns after
 * enqueueing; queries whose fast-pass result could not be verified are redone on
 * the exact path before return (that part synchronises the stream). */
cx_status cx_search_batch_device(cx_index* h, const float* d_queries, uint64_t B, uint64_t k,
                                 const cx_filter* filter, uint32_t* d_out_rows, float* d_out_score,
                                 float* d_out_distance, uint8_t* d_out_ids, uint32_t* d_out_n, void* stream);

/* Two-step form of cx_search_batch_device for callers that have more device work to enqueue behind
 * the search (the sharded exchange: pack -> all_gather -> merge).  _begin enqueues the scan and
 * returns without waiting; `stream` is ordered after the results, which are final unless _end says
 * otherwise.  _end waits, re-runs the queries whose fast-pass result could not be verified (rare) on
 * the tighter paths, and reports how many there were in *n_redone: if non-zero the output buffers
 * changed after `stream` may have consumed them and the caller must redo its dependent work.
 * *ticket may come back NULL (the call already ran to completion); _end then does nothing.  Every
 * _begin must be matched by one _end on the same thread before the next mutation of the index. */
cx_status cx_search_batch_device_begin(cx_index* h, const float* d_queries, uint64_t B, uint64_t k,
                                       const cx_filter* filter, uint32_t* d_out_rows, float* d_out_score,
                                       float* d_out_distance, uint8_t* d_out_ids, uint32_t* d_out_n, void* stream,
                                       void** ticket);
cx_status cx_search_batch_device_end(cx_index* h, void* ticket, uint64_t* n_redone);

/* The scan step of AutoLinker::run_cycle (linker/auto_linker.rs:215-264) for a batch of B new
 * nodes: search(embedding, k) (k = 100 in the reference, :221), skip the node itself
 * (:235-237, found through new_ids; NULL = none of them is in the index), keep
 * `score >= threshold` (SimilarityLinkRule, linker/rules.rs:50) in best-first order, at most
 * max_edges_per_node per node (:261).  Outputs are [B][max_edges_per_node]; out_n[b] = links
 * proposed for node b (to-id, score = edge weight).  Storage lookups, structural rules and
 * de-duplication against existing edges stay with the caller.  (Internally the search asks for
 * min(k, max_edges_per_node + 1) neighbours: the walk over the best-first list skips at most one entry --
 * the node itself -- and stops at the cap, so nothing further down can ever become a link.) */
cx_status cx_autolink_batch(cx_index* h, const uint8_t* new_ids, const float* embeddings, uint64_t B,
                            uint32_t len, uint64_t k, float threshold, uint32_t max_edges_per_node,
                            uint8_t* out_to_ids, float* out_score, uint32_t* out_n);
/* Device-resident form: embeddings [B][dim] f32 and optional d_self_rows [B] (0xFFFFFFFF = not
 * in the index) in HBM; d_scratch_* are [B][k] / [B] work buffers for the search results;
 * outputs [B][max_edges_per_node] (d_out_ids may be NULL). */
cx_status cx_autolink_batch_device(cx_index* h, const float* d_embeddings, uint64_t B, uint64_t k, float threshold,
                                   uint32_t max_edges_per_node, const uint32_t* d_self_rows,
                                   uint32_t* d_scratch_rows, float* d_scratch_score, float* d_scratch_distance,
                                   uint32_t* d_scratch_n, uint32_t* d_out_rows, float* d_out_score,
                                   uint8_t* d_out_ids, uint32_t* d_out_n, void* stream);

/* Row-sharded search (one process per GPU): the exchange step around the
 * caller's all-gather.  pack: a rank's device-resident local top-k -> payload of B*k*2 + 2 u64:
 * [B][k][2] slots (score-order key | distance bits, global row = row_offset + local row) and a
 * two-word trailer whose first word counts this rank's still unverified queries (d_ok = the
 * verification flags of a search in flight, cx_search_ticket_ok; NULL = all verified).
 * merge: the gathered payloads [world][B*k*2 + 2] -> global top-k per query in the single-index
 * order (score desc, NaN last, global row asc); *d_out_unverified (optional) = unverified queries
 * over all ranks: non-zero means some rank will change its list and the exchange must be repeated.
 * No reference counterpart: the reference is single-process (ARCHITECTURE.md:38). */
cx_status cx_pack_topk_device(const uint32_t* d_rows, const float* d_score, const float* d_distance,
                              const uint32_t* d_n, const uint32_t* d_ok, uint64_t B, uint64_t k,
                              uint64_t row_offset, uint64_t* d_payload, void* stream);
cx_status cx_merge_topk_device(const uint64_t* d_gathered, uint32_t world, uint64_t B, uint64_t k,
                               int64_t* d_out_rows, float* d_out_score, float* d_out_distance,
                               uint32_t* d_out_n, uint64_t* d_out_unverified, void* stream);
/* The auto-link candidate post-pass (linker/auto_linker.rs:224-264, rules.rs:42-62) on the MERGED lists of
 * a process-per-GPU sharded search (global rows): skip the node itself (d_self_rows [B], < 0 = not in the
 * index; may be NULL), keep score >= threshold, at most max_edges_per_node.  Outputs [B][max_edges_per_node]. */
cx_status cx_autolink_filter_device(const int64_t* d_rows, const float* d_score, const uint32_t* d_n,
                                    const int64_t* d_self_rows, uint64_t B, uint64_t k, float threshold,
                                    uint32_t max_edges_per_node, int64_t* d_out_rows, float* d_out_score,
                                    uint32_t* d_out_n, void* stream);
/* device pointer to the verification flags [B] of a search in flight (NULL ticket -> NULL) */
cx_status cx_search_ticket_ok(void* ticket, const uint32_t** d_ok);

/* VectorIndex::save / load, index.rs:437-472: bincode 1.3 layout of
 * (HashMap<Uuid,Vec<f32>>, HashMap<Uuid,NodeMetadata>, usize). */
cx_status cx_save(const cx_index* h, const char* path);
cx_status cx_load(const char* path, int device, cx_index** out);
cx_status cx_load_sharded(const char* path, const int* devices, uint32_t n_devices, cx_index** out);

/* id of shard-local row r (16 bytes) -- used when merging device results */
cx_status cx_row_id(const cx_index* h, uint32_t row, uint8_t out_id[16]);

/* ---- the callers' data formats either side of the scan ---------------------- */

/* What the node walk decided for a node of the redb `nodes` table. */
typedef enum cx_node_status {
  CX_NODE_OK = 0,               /* live node with an embedding of the right dimension: extracted / inserted */
  CX_NODE_NO_EMBEDDING = 1,     /* embedding: None (serve.rs:111 skips it) */
  CX_NODE_DELETED = 2,          /* tombstone: list_nodes(NodeFilter::new()) leaves it out (redb_storage.rs:347) */
  CX_NODE_DIM_MISMATCH = 3,     /* insert() is Err, the start-up loop moves on (serve.rs:112-114) */
  CX_NODE_NEEDS_HOST_DECODE = 4,/* non-empty `metadata` (bincode does not describe serde_json::Value, types.rs:142)
                                   or a timestamp form the walk does not parse: decode it with the caller's own
                                   deserializer and insert it through cx_insert */
  CX_NODE_CORRUPT = 5           /* the reference's deserializer fails too and skips the record (redb_storage.rs:707-710) */
} cx_node_status;

/* Embedding extractor for the start-up loop (serve.rs:105-123, api.rs:55-69): `values` holds the n raw values
 * of the `nodes` table back to back (bincode 1.3 of `Node`, types.rs:26-68; layout pinned by the golden bytes
 * of storage/redb_storage.rs:1834-1856), value i = values[offsets[i] .. offsets[i+1]).  One GPU thread per node
 * walks the variable-length fields; out_status[i] says what it found, out_ids [n][16], out_created_ns /
 * out_last_accessed_ns [n] (nanoseconds since the epoch), out_access_count [n] and out_rows [n][dim] (written
 * for CX_NODE_OK nodes only) are optional. */
cx_status cx_extract_embeddings(const uint8_t* values, const uint64_t* offsets, uint64_t n, uint32_t dim, int device,
                                uint8_t* out_ids, float* out_rows, int64_t* out_created_ns,
                                int64_t* out_last_accessed_ns, uint64_t* out_access_count, uint8_t* out_status);
/* The whole start-up loop in one call: every CX_NODE_OK node is inserted in the reference's order (newest
 * created_at first, ties in table order: redb_storage.rs:728); the embeddings move from the uploaded blob into
 * the store on the device.  out_status [n] and out_counts [6] (nodes per cx_node_status) are optional. */
cx_status cx_load_nodes(cx_index* h, const uint8_t* values, const uint64_t* offsets, uint64_t n, uint8_t* out_status,
                        uint64_t out_counts[6]);

/* ScoreDecayConfig (vector/scoring.rs:20-76) without the by_kind table: the caller resolves
 * `by_kind.get(kind).unwrap_or(daily_rate)` per candidate and passes it as kind_rate. */
typedef struct cx_decay_config {
  int32_t enabled;
  double max_age_days, min_factor, echo_weight, echo_cap;
} cx_decay_config;
/* apply_score_decay (vector/scoring.rs:84-114) for n search candidates at once, seg_len per query (0 = one
 * segment): idle_seconds = (now - last_accessed_at).num_seconds().  out_order (optional, [n]): per segment the
 * candidates' positions re-ranked by decayed score, best first, equal scores in candidate order -- the
 * re-rank of the search handler (http/routes.rs:945-949). */
cx_status cx_apply_score_decay(cx_index* h, const cx_decay_config* cfg, float recency_bias, uint64_t n, uint32_t seg_len,
                               const float* raw_score, const int64_t* idle_seconds, const uint64_t* access_count,
                               const double* kind_rate, float* out_score, uint32_t* out_order);

cx_status cx_get_stats(const cx_index* h, cx_stats* out);
/* Tuning / test hooks, all per index; set them while no search is running on the handle.
 * "force_path" 0 auto, 1 stream (K1), 2 tensor (K2), 3 exact; "tensor_min_batch" smallest query batch
 * the tensor pass serves (default 3); "tensor_phase_growth" growth factor of the scan phases (0/1 = one
 * phase, default auto); "tensor_sample_tiles" row tiles sampled for the cut-off bootstrap (0 = auto); "shadow" 0 = keep no bf16 copy (disables the tensor pass; before the first
 * insert); "profile" 1 = bracket the scan-pass kernels (bootstrap included) with CUDA events on their
 * stream (cx_get_stats: pass_kernel_ns); "blocking_sync" 1 = search calls sleep on an event instead of
 * spinning while the GPU works; "graphs" 0 = never replay repeated search shapes as a
 * CUDA graph; "tensor_pair" 1 = CTA-pair (cta_group::2) form of the tensor pass; "tensor_epi_warps"
 * 8 | 16; "stream_bf16" 0 = batches of up to four queries stream the fp32 rows instead of the
 * bf16 shadow (twice the bytes per pass; default 1 while the index keeps a shadow); "tensor_leftover_sms" 0 = leave the SMs idle that the (query tiles x row splits) grid of the tensor
 * pass does not cover (default 1: they take a share of the rows).  Every option leaves results identical.  (The result-corrupting measurement hook
 * "tensor_debug" exists only in -DCX_PROBE builds made by scripts/k2_probe.py, not in this library.) */
cx_status cx_set_option(cx_index* h, const char* key, int64_t value);

/* Test hook, needs no device: the tensor pass's launch plan -- queries per launch
 * group and row tiles (256 rows) per scan phase. */
cx_status cx_debug_tensor_plan(uint64_t n_queries, uint64_t n_rows, int sm_count, uint32_t sample_tiles,
                               uint32_t growth, uint32_t* groups, uint32_t* groups_n, uint32_t* phases,
                               uint32_t* phases_n, double* hits_per_kp);

const char* cx_last_error(void);
const char* cx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CORTEX_GPU_H */
