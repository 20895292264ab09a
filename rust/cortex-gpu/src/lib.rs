//! GpuVectorIndex: cortex_core::vector::VectorIndex over the B200 scan engine.
//!
//! Every trait method is one FFI call; any non-zero status becomes
//! CortexError::Validation(cx_last_error()) -- the only error variant the reference
//! index raises (crates/cortex-core/src/vector/index.rs:299-305, 438-459).
use cortex_core::error::{CortexError, Result};
use cortex_core::types::{Embedding, NodeId, NodeKind};
use cortex_core::vector::{SimilarityResult, VectorFilter, VectorIndex};
use cortex_gpu_sys as sys;
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::path::Path;
use std::ptr;

pub struct GpuVectorIndex {
    h: *mut sys::cx_index,
}
// search* are re-entrant in the library (per-call workspace + stream); mutators take &mut self.
unsafe impl Send for GpuVectorIndex {}
unsafe impl Sync for GpuVectorIndex {}

fn check(st: i32) -> Result<()> {
    if st == sys::CX_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::cx_last_error()) }.to_string_lossy().into_owned();
    Err(CortexError::Validation(msg))
}

struct CFilter {
    raw: sys::cx_filter,
    _kinds: Vec<CString>,
    _kind_ptrs: Vec<*const libc::c_char>,
    _excl: Vec<u8>,
    _agent: Option<CString>,
}

fn c_filter(f: &VectorFilter) -> CFilter {
    let kinds: Vec<CString> = f.kinds.iter().flatten().map(|k| CString::new(k.as_str()).unwrap()).collect();
    let kind_ptrs: Vec<_> = kinds.iter().map(|c| c.as_ptr()).collect();
    let excl: Vec<u8> = f.exclude.iter().flatten().flat_map(|id| id.as_bytes().to_vec()).collect();
    let agent = f.source_agent.as_ref().map(|a| CString::new(a.as_str()).unwrap());
    let raw = sys::cx_filter {
        has_kinds: f.kinds.is_some() as i32,
        kinds: kind_ptrs.as_ptr(),
        n_kinds: kind_ptrs.len() as u32,
        has_exclude: f.exclude.is_some() as i32,
        exclude_ids: excl.as_ptr(),
        n_exclude: (excl.len() / 16) as u32,
        has_source_agent: f.source_agent.is_some() as i32,
        source_agent: agent.as_ref().map_or(ptr::null(), |a| a.as_ptr()),
    };
    CFilter { raw, _kinds: kinds, _kind_ptrs: kind_ptrs, _excl: excl, _agent: agent }
}

impl GpuVectorIndex {
    /// The similarity step of DedupScanner::scan (linker/dedup.rs:65-127) as one self-join on the
    /// device: every unordered pair of live nodes with score >= threshold, once, as
    /// (node_a, node_b, similarity) with a inserted before b.  The scanner keeps its
    /// determine_action() loop (dedup.rs:130-171) and feeds it these pairs instead of calling
    /// search_threshold once per node.
    pub fn dedup_scan(&self, threshold: f32, per_node_cap: u32, max_pairs: usize)
        -> Result<(Vec<(NodeId, NodeId, f32)>, u64)> {
        let (mut a, mut b, mut sc) = (vec![0u8; 16 * max_pairs], vec![0u8; 16 * max_pairs], vec![0f32; max_pairs]);
        let (mut n, mut total) = (0u64, 0u64);
        check(unsafe {
            sys::cx_dedup_scan(self.h, threshold, per_node_cap, max_pairs as u64, a.as_mut_ptr(), b.as_mut_ptr(),
                               sc.as_mut_ptr(), &mut n, &mut total)
        })?;
        let id = |buf: &[u8], i: usize| NodeId::from_bytes(buf[16 * i..16 * i + 16].try_into().unwrap());
        Ok(((0..n as usize).map(|i| (id(&a, i), id(&b, i), sc[i])).collect(), total))
    }

    /// The scan step of AutoLinker::run_cycle (linker/auto_linker.rs:215-264) for a batch of new
    /// nodes: search(embedding, k), skip self, keep score >= threshold, at most max_edges per node.
    /// Returns, per node, the (neighbour, score) list in best-first order.
    pub fn autolink_batch(&self, nodes: &[(NodeId, Embedding)], k: usize, threshold: f32, max_edges: u32)
        -> Result<Vec<Vec<(NodeId, f32)>>> {
        if nodes.is_empty() {
            return Ok(Vec::new());
        }
        let (b, dim, me) = (nodes.len(), nodes[0].1.len(), max_edges as usize);
        if let Some((_, e)) = nodes.iter().find(|(_, e)| e.len() != dim) {
            // the flat buffer below has one stride: refuse mixed lengths instead of reading past it
            return Err(CortexError::Validation(format!(
                "Embedding dimension mismatch: expected {}, got {}", dim, e.len())));
        }
        let ids: Vec<u8> = nodes.iter().flat_map(|(id, _)| id.as_bytes().to_vec()).collect();
        let flat: Vec<f32> = nodes.iter().flat_map(|(_, e)| e.iter().copied()).collect();
        let (mut to, mut sc, mut n) = (vec![0u8; 16 * b * me], vec![0f32; b * me], vec![0u32; b]);
        check(unsafe {
            sys::cx_autolink_batch(self.h, ids.as_ptr(), flat.as_ptr(), b as u64, dim as u32, k as u64, threshold,
                                   max_edges, to.as_mut_ptr(), sc.as_mut_ptr(), n.as_mut_ptr())
        })?;
        Ok((0..b).map(|i| (0..n[i] as usize).map(|j| {
            let o = i * me + j;
            (NodeId::from_bytes(to[16 * o..16 * o + 16].try_into().unwrap()), sc[o])
        }).collect()).collect())
    }

    /// The start-up loop (serve.rs:105-123 / api.rs:55-69) in one call: `values` = the raw values of the
    /// redb `nodes` table back to back, `offsets[i]..offsets[i+1]` = value i.  Every live node with an
    /// embedding is inserted newest first; the returned status per node tells the caller which nodes
    /// (non-empty metadata) it has to decode itself (4 = CX_NODE_NEEDS_HOST_DECODE).
    pub fn load_nodes(&mut self, values: &[u8], offsets: &[u64]) -> Result<(Vec<u8>, [u64; 6])> {
        let n = offsets.len().saturating_sub(1);
        let (mut status, mut counts) = (vec![0u8; n], [0u64; 6]);
        check(unsafe {
            sys::cx_load_nodes(self.h, values.as_ptr(), offsets.as_ptr(), n as u64, status.as_mut_ptr(),
                               counts.as_mut_ptr())
        })?;
        Ok((status, counts))
    }

    /// apply_score_decay (vector/scoring.rs:84-114) for the candidates of one query, plus the re-rank of
    /// the search handler (http/routes.rs:945-949): returns (decayed scores, order best first).
    /// `idle_seconds[i]` = (now - node.last_accessed_at).num_seconds(), `kind_rate[i]` =
    /// config.by_kind.get(kind).unwrap_or(config.daily_rate).
    pub fn apply_score_decay(&self, cfg: &sys::cx_decay_config, recency_bias: f32, raw: &[f32], idle_seconds: &[i64],
                             access_count: &[u64], kind_rate: &[f64]) -> Result<(Vec<f32>, Vec<u32>)> {
        let n = raw.len();
        if idle_seconds.len() != n || access_count.len() != n || kind_rate.len() != n {
            return Err(CortexError::Validation("apply_score_decay: array lengths differ".into()));
        }
        let (mut out, mut order) = (vec![0f32; n], vec![0u32; n]);
        check(unsafe {
            sys::cx_apply_score_decay(self.h, cfg, recency_bias, n as u64, n as u32, raw.as_ptr(), idle_seconds.as_ptr(),
                                      access_count.as_ptr(), kind_rate.as_ptr(), out.as_mut_ptr(), order.as_mut_ptr())
        })?;
        Ok((out, order))
    }

    /// HnswIndex::new(dimension), index.rs:204-211
    pub fn new(dimension: usize) -> Result<Self> {
        Self::on_device(dimension, 0)
    }
    pub fn on_device(dimension: usize, device: i32) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { sys::cx_index_create(dimension as u32, device, &mut h) })?;
        Ok(Self { h })
    }
    /// One index row-sharded over several GPUs of the box, driven from this process: the same
    /// `VectorIndex` object (`Arc<RwLock<GpuVectorIndex>>`, serve.rs:101) then spans all of them.
    pub fn on_devices(dimension: usize, devices: &[i32]) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { sys::cx_index_create_sharded(dimension as u32, devices.as_ptr(), devices.len() as u32, &mut h) })?;
        Ok(Self { h })
    }
    pub fn load_on_devices(path: &Path, devices: &[i32]) -> Result<Self> {
        let p = CString::new(path.to_string_lossy().as_bytes()).unwrap();
        let mut h = ptr::null_mut();
        check(unsafe { sys::cx_load_sharded(p.as_ptr(), devices.as_ptr(), devices.len() as u32, &mut h) })?;
        Ok(Self { h })
    }
    pub fn shard_count(&self) -> usize {
        unsafe { sys::cx_shard_count(self.h) as usize }
    }
    /// Counters of the index (which pass served how many queries, fallbacks, bytes copied, store growth).
    pub fn stats(&self) -> Result<sys::cx_stats> {
        let mut st: sys::cx_stats = unsafe { std::mem::zeroed() };
        check(unsafe { sys::cx_get_stats(self.h, &mut st) })?;
        Ok(st)
    }
    /// Tuning hooks of cortex_gpu.h (`"graphs"`, `"stream_bf16"`, `"tensor_min_batch"`, ...); results are the
    /// same with every setting.  Call it while no search runs on the index (it takes `&mut self`).
    pub fn set_option(&mut self, key: &str, value: i64) -> Result<()> {
        let k = CString::new(key).unwrap();
        check(unsafe { sys::cx_set_option(self.h, k.as_ptr(), value) })
    }
    /// HnswIndex::set_metadata, index.rs:219-222
    pub fn set_metadata(&mut self, id: NodeId, kind: NodeKind, source_agent: String) {
        let k = CString::new(kind.as_str()).unwrap();
        let a = CString::new(source_agent).unwrap();
        unsafe { sys::cx_set_metadata(self.h, id.as_bytes().as_ptr(), k.as_ptr(), a.as_ptr()) };
    }
    /// Bulk form of the startup loop serve.rs:111-117
    pub fn insert_batch(&mut self, ids: &[NodeId], rows: &[f32], dim: usize) -> Result<()> {
        let flat: Vec<u8> = ids.iter().flat_map(|i| i.as_bytes().to_vec()).collect();
        check(unsafe { sys::cx_insert_batch(self.h, flat.as_ptr(), rows.as_ptr(), ids.len() as u64, dim as u32) })
    }
    fn collect(ids: &[u8], score: &[f32], dist: &[f32], n: usize) -> Vec<SimilarityResult> {
        (0..n)
            .map(|i| SimilarityResult {
                node_id: NodeId::from_slice(&ids[16 * i..16 * i + 16]).unwrap(),
                score: score[i],
                distance: dist[i],
            })
            .collect()
    }
}

impl Drop for GpuVectorIndex {
    fn drop(&mut self) {
        unsafe { sys::cx_index_destroy(self.h) }
    }
}

impl VectorIndex for GpuVectorIndex {
    fn insert(&mut self, id: NodeId, embedding: &Embedding) -> Result<()> {
        check(unsafe { sys::cx_insert(self.h, id.as_bytes().as_ptr(), embedding.as_ptr(), embedding.len() as u32) })
    }
    fn remove(&mut self, id: NodeId) -> Result<()> {
        check(unsafe { sys::cx_remove(self.h, id.as_bytes().as_ptr()) })
    }
    fn search(&self, query: &Embedding, k: usize, filter: Option<&VectorFilter>) -> Result<Vec<SimilarityResult>> {
        let kk = k.min(self.len()).max(1);
        let (mut ids, mut sc, mut di, mut n) = (vec![0u8; 16 * kk], vec![0f32; kk], vec![0f32; kk], 0u64);
        let cf = filter.map(c_filter);
        check(unsafe {
            sys::cx_search(self.h, query.as_ptr(), query.len() as u32, k.min(kk) as u64,
                           cf.as_ref().map_or(ptr::null(), |c| &c.raw), ids.as_mut_ptr(), sc.as_mut_ptr(),
                           di.as_mut_ptr(), &mut n)
        })?;
        Ok(Self::collect(&ids, &sc, &di, n as usize))
    }
    fn search_threshold(&self, query: &Embedding, threshold: f32, filter: Option<&VectorFilter>)
        -> Result<Vec<SimilarityResult>> {
        let cf = filter.map(c_filter);
        let mut cap = self.len().clamp(1, 4096);
        loop {
            let (mut ids, mut sc, mut di) = (vec![0u8; 16 * cap], vec![0f32; cap], vec![0f32; cap]);
            let (mut n, mut total) = (0u64, 0u64);
            check(unsafe {
                sys::cx_search_threshold(self.h, query.as_ptr(), query.len() as u32, threshold,
                                         cf.as_ref().map_or(ptr::null(), |c| &c.raw), cap as u64,
                                         ids.as_mut_ptr(), sc.as_mut_ptr(), di.as_mut_ptr(), &mut n, &mut total)
            })?;
            if total as usize <= cap {
                return Ok(Self::collect(&ids, &sc, &di, n as usize));
            }
            cap = total as usize;
        }
    }
    fn search_batch(&self, queries: &[(NodeId, Embedding)], k: usize, filter: Option<&VectorFilter>)
        -> Result<HashMap<NodeId, Vec<SimilarityResult>>> {
        if queries.is_empty() {
            return Ok(HashMap::new());
        }
        // the reference searches every query on its own (index.rs:397-403), so lengths may differ:
        // one call per distinct length (normally exactly one), each over a flat buffer of that stride
        let k = k.min(self.len()).max(1);
        let mut by_len: HashMap<usize, Vec<usize>> = HashMap::new();
        for (i, (_, e)) in queries.iter().enumerate() {
            by_len.entry(e.len()).or_default().push(i);
        }
        let cf = filter.map(c_filter);
        let mut map = HashMap::with_capacity(queries.len());
        for (dim, idx) in by_len {
            let b = idx.len();
            let flat: Vec<f32> = idx.iter().flat_map(|&i| queries[i].1.iter().copied()).collect();
            let (mut ids, mut sc, mut di, mut n) =
                (vec![0u8; 16 * b * k], vec![0f32; b * k], vec![0f32; b * k], vec![0u64; b]);
            check(unsafe {
                sys::cx_search_batch(self.h, flat.as_ptr(), b as u64, dim as u32, k as u64,
                                     cf.as_ref().map_or(ptr::null(), |c| &c.raw), ids.as_mut_ptr(),
                                     sc.as_mut_ptr(), di.as_mut_ptr(), n.as_mut_ptr())
            })?;
            for (j, &i) in idx.iter().enumerate() {
                let r = Self::collect(&ids[16 * j * k..], &sc[j * k..], &di[j * k..], n[j] as usize);
                map.insert(queries[i].0, r);
            }
        }
        Ok(map)
    }
    fn len(&self) -> usize {
        unsafe { sys::cx_len(self.h) as usize }
    }
    fn rebuild(&mut self) -> Result<()> {
        check(unsafe { sys::cx_rebuild(self.h) })
    }
    fn save(&self, path: &Path) -> Result<()> {
        let p = CString::new(path.to_string_lossy().as_bytes()).unwrap();
        check(unsafe { sys::cx_save(self.h, p.as_ptr()) })
    }
    fn load(path: &Path) -> Result<Self> {
        let p = CString::new(path.to_string_lossy().as_bytes()).unwrap();
        let mut h = ptr::null_mut();
        check(unsafe { sys::cx_load(p.as_ptr(), 0, &mut h) })?;
        Ok(Self { h })
    }
}
