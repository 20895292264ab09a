//! Raw bindings of include/cortex_gpu.h, one declaration per exported symbol.
#![allow(non_camel_case_types)]
use libc::{c_char, c_int, c_void};

#[repr(C)]
pub struct cx_index { _private: [u8; 0] }

#[repr(C)]
pub struct cx_filter {
    pub has_kinds: i32,
    pub kinds: *const *const c_char,
    pub n_kinds: u32,
    pub has_exclude: i32,
    pub exclude_ids: *const u8,
    pub n_exclude: u32,
    pub has_source_agent: i32,
    pub source_agent: *const c_char,
}

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct cx_stats {
    pub kernel_launches: u64,
    pub queries_stream: u64,
    pub queries_tensor: u64,
    pub queries_exact: u64,
    pub fallbacks: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub pass_kernel_ns: u64,
    pub pass_kernel_launches: u64,
    pub graph_launches: u64,
    pub grow_events: u64,
    pub irregular_rows: u64,
    pub capacity_rows: u64,
    pub in_place_growth: u64,
    pub grow_ns: u64,
    pub grow_ns_max: u64,
    pub grow_waits: u64,
    pub unverified_overflow: u64,
    pub unverified_near_ties: u64,
    pub unverified_other: u64,
    pub queries_stream_bf16: u64,
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct cx_decay_config {
    pub enabled: i32,
    pub max_age_days: f64,
    pub min_factor: f64,
    pub echo_weight: f64,
    pub echo_cap: f64,
}

pub const CX_OK: c_int = 0;

extern "C" {
    pub fn cx_index_create(dimension: u32, device: c_int, out: *mut *mut cx_index) -> c_int;
    pub fn cx_index_create_sharded(dimension: u32, devices: *const c_int, n_devices: u32, out: *mut *mut cx_index) -> c_int;
    pub fn cx_shard_count(h: *const cx_index) -> u32;
    pub fn cx_index_destroy(h: *mut cx_index);
    pub fn cx_insert(h: *mut cx_index, id: *const u8, embedding: *const f32, len: u32) -> c_int;
    pub fn cx_insert_batch(h: *mut cx_index, ids: *const u8, rows: *const f32, n: u64, len: u32) -> c_int;
    pub fn cx_insert_batch_device(h: *mut cx_index, ids: *const u8, d_rows: *const f32, n: u64, len: u32) -> c_int;
    pub fn cx_remove(h: *mut cx_index, id: *const u8) -> c_int;
    pub fn cx_set_metadata(h: *mut cx_index, id: *const u8, kind: *const c_char, agent: *const c_char) -> c_int;
    pub fn cx_len(h: *const cx_index) -> u64;
    pub fn cx_dimension(h: *const cx_index) -> u32;
    pub fn cx_rebuild(h: *mut cx_index) -> c_int;
    pub fn cx_reserve(h: *mut cx_index, n_rows: u64) -> c_int;
    pub fn cx_search(h: *mut cx_index, query: *const f32, qlen: u32, k: u64, filter: *const cx_filter,
                     out_ids: *mut u8, out_score: *mut f32, out_distance: *mut f32, out_n: *mut u64) -> c_int;
    pub fn cx_search_threshold(h: *mut cx_index, query: *const f32, qlen: u32, threshold: f32,
                               filter: *const cx_filter, cap: u64, out_ids: *mut u8, out_score: *mut f32,
                               out_distance: *mut f32, out_n: *mut u64, out_total: *mut u64) -> c_int;
    pub fn cx_search_threshold_batch(h: *mut cx_index, queries: *const f32, b: u64, qlen: u32, threshold: f32,
                                     filter: *const cx_filter, cap: u64, out_ids: *mut u8, out_score: *mut f32,
                                     out_distance: *mut f32, out_n: *mut u64, out_total: *mut u64) -> c_int;
    pub fn cx_dedup_scan(h: *mut cx_index, threshold: f32, per_node_cap: u32, max_pairs: u64, out_a_ids: *mut u8,
                         out_b_ids: *mut u8, out_score: *mut f32, out_n: *mut u64, out_total: *mut u64) -> c_int;
    pub fn cx_autolink_batch(h: *mut cx_index, new_ids: *const u8, embeddings: *const f32, b: u64, len: u32, k: u64,
                             threshold: f32, max_edges_per_node: u32, out_to_ids: *mut u8, out_score: *mut f32,
                             out_n: *mut u32) -> c_int;
    pub fn cx_autolink_batch_device(h: *mut cx_index, d_embeddings: *const f32, b: u64, k: u64, threshold: f32,
                                    max_edges_per_node: u32, d_self_rows: *const u32, d_scratch_rows: *mut u32,
                                    d_scratch_score: *mut f32, d_scratch_distance: *mut f32, d_scratch_n: *mut u32,
                                    d_out_rows: *mut u32, d_out_score: *mut f32, d_out_ids: *mut u8,
                                    d_out_n: *mut u32, stream: *mut c_void) -> c_int;
    pub fn cx_search_batch(h: *mut cx_index, queries: *const f32, b: u64, qlen: u32, k: u64,
                           filter: *const cx_filter, out_ids: *mut u8, out_score: *mut f32,
                           out_distance: *mut f32, out_n: *mut u64) -> c_int;
    pub fn cx_search_batch_device(h: *mut cx_index, d_queries: *const f32, b: u64, k: u64,
                                  filter: *const cx_filter, d_out_rows: *mut u32, d_out_score: *mut f32,
                                  d_out_distance: *mut f32, d_out_ids: *mut u8, d_out_n: *mut u32,
                                  stream: *mut c_void) -> c_int;
    pub fn cx_search_batch_device_begin(h: *mut cx_index, d_queries: *const f32, b: u64, k: u64,
                                        filter: *const cx_filter, d_out_rows: *mut u32, d_out_score: *mut f32,
                                        d_out_distance: *mut f32, d_out_ids: *mut u8, d_out_n: *mut u32,
                                        stream: *mut c_void, ticket: *mut *mut c_void) -> c_int;
    pub fn cx_search_batch_device_end(h: *mut cx_index, ticket: *mut c_void, n_redone: *mut u64) -> c_int;
    pub fn cx_search_ticket_ok(ticket: *mut c_void, d_ok: *mut *const u32) -> c_int;
    pub fn cx_pack_topk_device(d_rows: *const u32, d_score: *const f32, d_distance: *const f32, d_n: *const u32,
                               d_ok: *const u32, b: u64, k: u64, row_offset: u64, d_payload: *mut u64,
                               stream: *mut c_void) -> c_int;
    pub fn cx_merge_topk_device(d_gathered: *const u64, world: u32, b: u64, k: u64, d_out_rows: *mut i64,
                                d_out_score: *mut f32, d_out_distance: *mut f32, d_out_n: *mut u32,
                                d_out_unverified: *mut u64, stream: *mut c_void) -> c_int;
    pub fn cx_autolink_filter_device(d_rows: *const i64, d_score: *const f32, d_n: *const u32, d_self_rows: *const i64,
                                     b: u64, k: u64, threshold: f32, max_edges_per_node: u32, d_out_rows: *mut i64,
                                     d_out_score: *mut f32, d_out_n: *mut u32, stream: *mut c_void) -> c_int;
    pub fn cx_debug_tensor_plan(n_queries: u64, n_rows: u64, sm_count: c_int, sample_tiles: u32, growth: u32,
                                groups: *mut u32, groups_n: *mut u32, phases: *mut u32, phases_n: *mut u32,
                                hits_per_kp: *mut f64) -> c_int;
    pub fn cx_save(h: *const cx_index, path: *const c_char) -> c_int;
    pub fn cx_load(path: *const c_char, device: c_int, out: *mut *mut cx_index) -> c_int;
    pub fn cx_load_sharded(path: *const c_char, devices: *const c_int, n_devices: u32, out: *mut *mut cx_index) -> c_int;
    pub fn cx_row_id(h: *const cx_index, row: u32, out_id: *mut u8) -> c_int;
    pub fn cx_extract_embeddings(values: *const u8, offsets: *const u64, n: u64, dim: u32, device: c_int,
                                 out_ids: *mut u8, out_rows: *mut f32, out_created_ns: *mut i64,
                                 out_last_accessed_ns: *mut i64, out_access_count: *mut u64, out_status: *mut u8) -> c_int;
    pub fn cx_load_nodes(h: *mut cx_index, values: *const u8, offsets: *const u64, n: u64, out_status: *mut u8,
                         out_counts: *mut u64) -> c_int;
    pub fn cx_apply_score_decay(h: *mut cx_index, cfg: *const cx_decay_config, recency_bias: f32, n: u64, seg_len: u32,
                                raw_score: *const f32, idle_seconds: *const i64, access_count: *const u64,
                                kind_rate: *const f64, out_score: *mut f32, out_order: *mut u32) -> c_int;
    pub fn cx_get_stats(h: *const cx_index, out: *mut cx_stats) -> c_int;
    pub fn cx_set_option(h: *mut cx_index, key: *const c_char, value: i64) -> c_int;
    pub fn cx_last_error() -> *const c_char;
    pub fn cx_version() -> *const c_char;
}
