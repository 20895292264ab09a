"""Host-side mirror of cortex-core's vector-index surface over the C ABI.

Names, argument meaning and error behaviour follow
/root/reference/crates/cortex-core/src/vector/index.rs:
    SimilarityResult   :9-15      VectorFilter      :17-47
    trait VectorIndex  :50-99     HnswIndex::new    :204-211
    set_metadata       :219-222
`GpuVectorIndex` is what `impl VectorIndex for GpuVectorIndex` looks like from
Python; every method is one call into libcortex_gpu.so.  There is no fallback:
if the library or a CUDA device is missing the constructor raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi


class CortexError(Exception):
    """cortex_core::CortexError (error.rs:7-50).  The index only ever raises the
    Validation variant (index.rs:299-305, 438-459)."""

    def __init__(self, message: str, status: int = _capi.CX_ERR_VALIDATION):
        super().__init__(f"Validation error: {message}")
        self.message = message
        self.status = status


def _check(st: int) -> None:
    if st != _capi.CX_OK:
        raise CortexError(_capi.load().cx_last_error().decode(errors="replace"), st)


def validate_kind(kind: str) -> str:
    """NodeKind::new (types.rs:76-90): non-empty, lowercase alphanumeric + hyphens."""
    if not kind:
        raise CortexError("NodeKind cannot be empty")
    if not all((c.isascii() and (c.islower() or c.isdigit())) or c == "-" for c in kind):
        raise CortexError(f"NodeKind '{kind}' must be lowercase alphanumeric + hyphens only")
    return kind


@dataclass
class SimilarityResult:
    node_id: bytes      # 16 raw uuid bytes
    score: float        # cosine similarity clamped to [0, 1]
    distance: float     # 1 - similarity, unclamped


@dataclass
class VectorFilter:
    kinds: Optional[List[str]] = None
    exclude: Optional[List[bytes]] = None
    source_agent: Optional[str] = None

    @classmethod
    def new(cls) -> "VectorFilter":
        return cls()

    def with_kinds(self, kinds: Sequence[str]) -> "VectorFilter":
        self.kinds = [validate_kind(k) for k in kinds]
        return self

    def excluding(self, ids: Sequence[bytes]) -> "VectorFilter":
        self.exclude = [bytes(i) for i in ids]
        return self

    def with_source_agent(self, agent: str) -> "VectorFilter":
        self.source_agent = agent
        return self


@dataclass
class _CFilter:
    struct: _capi.CxFilter
    keep: list = field(default_factory=list)

    @property
    def ptr(self):
        return C.addressof(self.struct)


def _c_filter(f: Optional[VectorFilter]) -> Optional[_CFilter]:
    if f is None:
        return None
    cf = _CFilter(_capi.CxFilter())
    s = cf.struct
    if f.kinds is not None:
        arr = (C.c_char_p * max(1, len(f.kinds)))(*[k.encode() for k in f.kinds])
        cf.keep.append(arr)
        s.has_kinds, s.kinds, s.n_kinds = 1, C.cast(arr, C.POINTER(C.c_char_p)), len(f.kinds)
    if f.exclude is not None:
        ex = np.frombuffer(b"".join(bytes(e) for e in f.exclude), dtype=np.uint8).copy()
        if ex.size != 16 * len(f.exclude):
            raise CortexError("exclude ids must be 16 bytes each")
        cf.keep.append(ex)
        s.has_exclude, s.exclude_ids, s.n_exclude = 1, (ex.ctypes.data if ex.size else None), len(f.exclude)
    if f.source_agent is not None:
        b = f.source_agent.encode()
        cf.keep.append(b)
        s.has_source_agent, s.source_agent = 1, b
    return cf


def _id16(node_id) -> np.ndarray:
    a = np.frombuffer(bytes(node_id), dtype=np.uint8)
    if a.size != 16:
        raise CortexError("node id must be 16 bytes")
    return a.copy()


def extract_embeddings(blob: np.ndarray, offsets: np.ndarray, dim: int, device: int = 0):
    """cx_extract_embeddings: walk the raw `nodes` table values on the GPU.  Returns a dict with status [n],
    ids [n,16], rows [n,dim] (valid where status == 0), created_ns, last_accessed_ns, access_count."""
    L = _capi.load()
    blob = np.ascontiguousarray(blob, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = offsets.size - 1
    out = {"status": np.zeros(n, np.uint8), "ids": np.zeros((n, 16), np.uint8), "rows": np.zeros((n, dim), np.float32),
           "created_ns": np.zeros(n, np.int64), "last_accessed_ns": np.zeros(n, np.int64),
           "access_count": np.zeros(n, np.uint64)}
    _check(L.cx_extract_embeddings(blob.ctypes.data, offsets.ctypes.data, n, dim, device, out["ids"].ctypes.data,
                                   out["rows"].ctypes.data, out["created_ns"].ctypes.data,
                                   out["last_accessed_ns"].ctypes.data, out["access_count"].ctypes.data,
                                   out["status"].ctypes.data))
    return out


class GpuVectorIndex:
    """B200-resident exact-scan index behind the reference's VectorIndex surface."""

    def __init__(self, dimension: int, device: int = 0, _handle: Optional[int] = None,
                 devices: Optional[Sequence[int]] = None):
        """device: one CUDA ordinal.  devices: a list of ordinals -- the index is row-sharded over them and
        driven from this process (cx_index_create_sharded); queries / results of the device-resident calls
        then live on devices[0]."""
        self._L = _capi.load()
        self.dimension = int(dimension)
        self.devices = list(devices) if devices is not None else None
        self.device = self.devices[0] if self.devices else device
        if _handle is not None:
            self._h = C.c_void_p(_handle)
        elif self.devices is not None:
            h = C.c_void_p()
            arr = (C.c_int * len(self.devices))(*self.devices)
            _check(self._L.cx_index_create_sharded(self.dimension, arr, len(self.devices), C.byref(h)))
            self._h = h
        else:
            h = C.c_void_p()
            _check(self._L.cx_index_create(self.dimension, device, C.byref(h)))
            self._h = h

    @property
    def shard_count(self) -> int:
        return int(self._L.cx_shard_count(self._h))

    # HnswIndex::with_metadata (index.rs:214-216) is an alias of new()
    @classmethod
    def with_metadata(cls, dimension: int, device: int = 0) -> "GpuVectorIndex":
        return cls(dimension, device)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.cx_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- mutation -----------------------------------------------------------
    def insert(self, node_id, embedding) -> None:
        v = np.ascontiguousarray(embedding, dtype=np.float32).reshape(-1)
        i = _id16(node_id)
        _check(self._L.cx_insert(self._h, i.ctypes.data, v.ctypes.data, v.size))

    def insert_batch(self, ids: np.ndarray, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.uint8).reshape(-1, 16)
        if rows.ndim != 2 or ids.shape[0] != rows.shape[0]:
            raise CortexError("insert_batch needs ids [n,16] and rows [n,dim]")
        _check(self._L.cx_insert_batch(self._h, ids.ctypes.data, rows.ctypes.data, rows.shape[0], rows.shape[1]))

    def insert_batch_device(self, ids: np.ndarray, d_rows) -> None:
        """Append rows that already live in HBM (torch.float32 CUDA tensor [n, dim]); ids on the host."""
        ids = np.ascontiguousarray(ids, dtype=np.uint8).reshape(-1, 16)
        assert d_rows.is_cuda and d_rows.is_contiguous() and d_rows.shape[0] == ids.shape[0]
        _check(self._L.cx_insert_batch_device(self._h, ids.ctypes.data, d_rows.data_ptr(), d_rows.shape[0],
                                              d_rows.shape[1]))

    def remove(self, node_id) -> None:
        _check(self._L.cx_remove(self._h, _id16(node_id).ctypes.data))

    def set_metadata(self, node_id, kind: str, source_agent: str) -> None:
        validate_kind(kind)
        _check(self._L.cx_set_metadata(self._h, _id16(node_id).ctypes.data, kind.encode(), source_agent.encode()))

    def reserve(self, n_rows: int) -> None:
        _check(self._L.cx_reserve(self._h, n_rows))

    def rebuild(self) -> None:
        _check(self._L.cx_rebuild(self._h))

    # ---- queries ------------------------------------------------------------
    def __len__(self) -> int:
        return int(self._L.cx_len(self._h))

    def len(self) -> int:
        return len(self)

    def is_empty(self) -> bool:
        return len(self) == 0

    def search(self, query, k: int, filter: Optional[VectorFilter] = None) -> List[SimilarityResult]:
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        kk = max(1, min(int(k), max(1, len(self))))
        ids = np.zeros((kk, 16), np.uint8)
        sc = np.zeros(kk, np.float32)
        di = np.zeros(kk, np.float32)
        n = C.c_uint64(0)
        cf = _c_filter(filter)
        _check(self._L.cx_search(self._h, q.ctypes.data, q.size, min(int(k), kk), cf.ptr if cf else None,
                                 ids.ctypes.data, sc.ctypes.data, di.ctypes.data, C.byref(n)))
        return [SimilarityResult(ids[i].tobytes(), float(sc[i]), float(di[i])) for i in range(n.value)]

    def search_threshold(self, query, threshold: float,
                         filter: Optional[VectorFilter] = None) -> List[SimilarityResult]:
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        cf = _c_filter(filter)
        cap = min(max(1, len(self)), 4096)
        while True:
            ids = np.zeros((cap, 16), np.uint8)
            sc = np.zeros(cap, np.float32)
            di = np.zeros(cap, np.float32)
            n, total = C.c_uint64(0), C.c_uint64(0)
            _check(self._L.cx_search_threshold(self._h, q.ctypes.data, q.size, C.c_float(threshold),
                                               cf.ptr if cf else None, cap, ids.ctypes.data, sc.ctypes.data,
                                               di.ctypes.data, C.byref(n), C.byref(total)))
            if total.value <= cap:
                break
            cap = int(total.value)
        return [SimilarityResult(ids[i].tobytes(), float(sc[i]), float(di[i])) for i in range(n.value)]

    def search_threshold_batch_arrays(self, queries: np.ndarray, threshold: float, cap: int,
                                      filter: Optional[VectorFilter] = None):
        """search_threshold for B queries at once.  Returns ids [B,cap,16], score [B,cap],
        distance [B,cap], n [B] (entries written), total [B] (rows that qualify)."""
        Q = np.ascontiguousarray(queries, dtype=np.float32)
        if Q.ndim != 2:
            raise CortexError("queries must be [B, dim]")
        B, qlen = Q.shape
        cap = max(1, int(cap))
        ids = np.zeros((B, cap, 16), np.uint8)
        sc = np.zeros((B, cap), np.float32)
        di = np.zeros((B, cap), np.float32)
        n = np.zeros(B, np.uint64)
        total = np.zeros(B, np.uint64)
        cf = _c_filter(filter)
        _check(self._L.cx_search_threshold_batch(self._h, Q.ctypes.data, B, qlen, C.c_float(threshold),
                                                 cf.ptr if cf else None, cap, ids.ctypes.data, sc.ctypes.data,
                                                 di.ctypes.data, n.ctypes.data, total.ctypes.data))
        return ids, sc, di, n, total

    def dedup_scan(self, threshold: float, per_node_cap: int = 64, max_pairs: int = 1 << 20):
        """DedupScanner::scan's similarity step (linker/dedup.rs:65-127) as one self-join.
        Returns (a_ids [P,16], b_ids [P,16], score [P], total) -- every unordered pair once,
        a inserted before b, ordered by (a, score desc, b)."""
        a = np.zeros((max_pairs, 16), np.uint8)
        b = np.zeros((max_pairs, 16), np.uint8)
        sc = np.zeros(max_pairs, np.float32)
        n, total = C.c_uint64(0), C.c_uint64(0)
        _check(self._L.cx_dedup_scan(self._h, C.c_float(threshold), int(per_node_cap), int(max_pairs), a.ctypes.data,
                                     b.ctypes.data, sc.ctypes.data, C.byref(n), C.byref(total)))
        return a[:n.value], b[:n.value], sc[:n.value], int(total.value)

    def search_batch_arrays(self, queries: np.ndarray, k: int, filter: Optional[VectorFilter] = None
                            ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Array form: ids [B,k,16], score [B,k], distance [B,k], n [B]."""
        Q = np.ascontiguousarray(queries, dtype=np.float32)
        if Q.ndim != 2:
            raise CortexError("queries must be [B, dim]")
        B, qlen = Q.shape
        kk = max(1, int(k))
        ids = np.zeros((B, kk, 16), np.uint8)
        sc = np.zeros((B, kk), np.float32)
        di = np.zeros((B, kk), np.float32)
        n = np.zeros(B, np.uint64)
        cf = _c_filter(filter)
        _check(self._L.cx_search_batch(self._h, Q.ctypes.data, B, qlen, int(k), cf.ptr if cf else None,
                                       ids.ctypes.data, sc.ctypes.data, di.ctypes.data, n.ctypes.data))
        return ids, sc, di, n

    def search_batch(self, queries: Iterable[Tuple[bytes, Sequence[float]]], k: int,
                     filter: Optional[VectorFilter] = None) -> Dict[bytes, List[SimilarityResult]]:
        """index.rs:390-410: map keyed by the query's node id."""
        queries = list(queries)
        if not queries:
            return {}
        Q = np.stack([np.asarray(e, dtype=np.float32) for _, e in queries])
        ids, sc, di, n = self.search_batch_arrays(Q, k, filter)
        out: Dict[bytes, List[SimilarityResult]] = {}
        for b, (qid, _) in enumerate(queries):
            out[bytes(qid)] = [SimilarityResult(ids[b, i].tobytes(), float(sc[b, i]), float(di[b, i]))
                               for i in range(int(n[b]))]
        return out

    def search_batch_device(self, d_queries, k: int, filter: Optional[VectorFilter] = None, stream: int = 0,
                            out=None, ids_out=None):
        """Queries and results stay in HBM.  d_queries: torch.float32 CUDA tensor [B, dim].
        Returns (rows int32 [B,k], score [B,k], distance [B,k], n int32 [B]) CUDA tensors.
        ids_out (optional): uint8 CUDA tensor [B,k,16] that receives the node ids."""
        import torch

        assert d_queries.is_cuda and d_queries.dtype == torch.float32 and d_queries.is_contiguous()
        B = d_queries.shape[0]
        if out is None:
            dev = d_queries.device
            out = (torch.empty((B, k), dtype=torch.int32, device=dev),
                   torch.empty((B, k), dtype=torch.float32, device=dev),
                   torch.empty((B, k), dtype=torch.float32, device=dev),
                   torch.empty((B,), dtype=torch.int32, device=dev))
        rows, sc, di, n = out
        cf = _c_filter(filter)
        _check(self._L.cx_search_batch_device(self._h, d_queries.data_ptr(), B, int(k), cf.ptr if cf else None,
                                              rows.data_ptr(), sc.data_ptr(), di.data_ptr(),
                                              ids_out.data_ptr() if ids_out is not None else None,
                                              n.data_ptr(), C.c_void_p(stream)))
        return out

    def search_batch_device_begin(self, d_queries, k: int, filter: Optional[VectorFilter] = None, stream: int = 0,
                                  out=None, ids_out=None):
        """Enqueue a device-resident search without waiting (cx_search_batch_device_begin).  Returns
        (out, ticket): `stream` is ordered after the results; call search_batch_device_end(ticket)
        before the next mutation -- it reports how many queries had to be redone."""
        import torch

        assert d_queries.is_cuda and d_queries.dtype == torch.float32 and d_queries.is_contiguous()
        B = d_queries.shape[0]
        if out is None:
            dev = d_queries.device
            out = (torch.empty((B, k), dtype=torch.int32, device=dev),
                   torch.empty((B, k), dtype=torch.float32, device=dev),
                   torch.empty((B, k), dtype=torch.float32, device=dev),
                   torch.empty((B,), dtype=torch.int32, device=dev))
        rows, sc, di, n = out
        cf = _c_filter(filter)
        ticket = C.c_void_p()
        _check(self._L.cx_search_batch_device_begin(self._h, d_queries.data_ptr(), B, int(k), cf.ptr if cf else None,
                                                    rows.data_ptr(), sc.data_ptr(), di.data_ptr(),
                                                    ids_out.data_ptr() if ids_out is not None else None,
                                                    n.data_ptr(), C.c_void_p(stream), C.byref(ticket)))
        return out, ticket

    def ticket_ok_ptr(self, ticket) -> int:
        """Device address of the verification flags of a search in flight (0 = none: all verified)."""
        p = C.c_void_p()
        _check(self._L.cx_search_ticket_ok(ticket, C.byref(p)))
        return p.value or 0

    def ticket_ok(self, ticket, B: int):
        """The verification flags as a borrowed int32 CUDA tensor [B] (None if there are none)."""
        import torch

        addr = self.ticket_ok_ptr(ticket)
        if not addr:
            return None

        class _Borrowed:
            __cuda_array_interface__ = {"shape": (B,), "typestr": "<i4", "data": (addr, False), "version": 3}

        return torch.as_tensor(_Borrowed(), device=f"cuda:{self.device}")

    def search_batch_device_end(self, ticket) -> int:
        n = C.c_uint64(0)
        _check(self._L.cx_search_batch_device_end(self._h, ticket, C.byref(n)))
        return int(n.value)

    def autolink_batch(self, new_nodes: Iterable[Tuple[bytes, Sequence[float]]], threshold: float = 0.75,
                       k: int = 100, max_edges_per_node: int = 50) -> Dict[bytes, List[Tuple[bytes, float]]]:
        """The scan step of AutoLinker::run_cycle (linker/auto_linker.rs:215-264) for a batch of new
        nodes: {node id: [(neighbour id, score)...]} with score >= threshold, best first, self
        skipped, at most max_edges_per_node each (defaults: linker/config.rs:59-66, auto_linker.rs:221)."""
        new_nodes = list(new_nodes)
        if not new_nodes:
            return {}
        E = np.ascontiguousarray(np.stack([np.asarray(e, dtype=np.float32) for _, e in new_nodes]))
        ids = np.ascontiguousarray(np.stack([_id16(i) for i, _ in new_nodes]))
        B, me = E.shape[0], int(max_edges_per_node)
        to = np.zeros((B, me, 16), np.uint8)
        sc = np.zeros((B, me), np.float32)
        n = np.zeros(B, np.uint32)
        _check(self._L.cx_autolink_batch(self._h, ids.ctypes.data, E.ctypes.data, B, E.shape[1], int(k),
                                         C.c_float(threshold), me, to.ctypes.data, sc.ctypes.data, n.ctypes.data))
        return {bytes(nid): [(to[b, j].tobytes(), float(sc[b, j])) for j in range(int(n[b]))]
                for b, (nid, _) in enumerate(new_nodes)}

    def autolink_batch_device(self, d_embeddings, k: int = 100, threshold: float = 0.75,
                              max_edges_per_node: int = 50, d_self_rows=None, stream: int = 0, bufs=None):
        """cx_autolink_batch_device: the auto-link scan step with new-node embeddings [B, dim] and results in
        HBM (single-device index).  Returns ((rows int32 [B,me], score [B,me], n int32 [B]), bufs); pass
        `bufs` back in to reuse the scratch and output buffers."""
        import torch

        assert d_embeddings.is_cuda and d_embeddings.dtype == torch.float32 and d_embeddings.is_contiguous()
        B, me, dev = d_embeddings.shape[0], int(max_edges_per_node), d_embeddings.device
        kk = min(int(k), max(1, len(self)))
        if bufs is None:
            bufs = (torch.empty((B, kk), dtype=torch.int32, device=dev), torch.empty((B, kk), dtype=torch.float32, device=dev),
                    torch.empty((B, kk), dtype=torch.float32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
                    torch.empty((B, me), dtype=torch.int32, device=dev), torch.empty((B, me), dtype=torch.float32, device=dev),
                    torch.empty((B,), dtype=torch.int32, device=dev))
        sr, ss, sd, sn, orow, osc, on = bufs
        _check(self._L.cx_autolink_batch_device(self._h, d_embeddings.data_ptr(), B, kk, C.c_float(threshold), me,
                                                d_self_rows.data_ptr() if d_self_rows is not None else None,
                                                sr.data_ptr(), ss.data_ptr(), sd.data_ptr(), sn.data_ptr(),
                                                orow.data_ptr(), osc.data_ptr(), None, on.data_ptr(), C.c_void_p(stream)))
        return (orow, osc, on), bufs

    # ---- the callers' data formats either side of the scan -------------------------
    def load_nodes(self, blob: np.ndarray, offsets: np.ndarray):
        """The start-up loop (serve.rs:105-123) in one call: raw values of the redb `nodes` table in, every live
        node with an embedding inserted newest first.  Returns (status uint8 [n], counts [6])."""
        blob = np.ascontiguousarray(blob, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        status = np.zeros(max(1, n), np.uint8)
        counts = np.zeros(6, np.uint64)
        _check(self._L.cx_load_nodes(self._h, blob.ctypes.data, offsets.ctypes.data, n, status.ctypes.data,
                                     counts.ctypes.data))
        return status[:n], counts

    def apply_score_decay(self, raw_score, idle_seconds, access_count, kind_rate, recency_bias: float = 0.15,
                          seg_len: int = 0, enabled: bool = True, max_age_days: float = 365.0, min_factor: float = 0.1,
                          echo_weight: float = 0.05, echo_cap: float = 2.0, rerank: bool = False):
        """apply_score_decay (vector/scoring.rs:84-114) for n candidates (defaults = ScoreDecayConfig::default()).
        Returns the decayed scores and, with rerank=True, the per-segment order (routes.rs:945-949)."""
        raw = np.ascontiguousarray(raw_score, np.float32)
        idle = np.ascontiguousarray(idle_seconds, np.int64)
        acc = np.ascontiguousarray(access_count, np.uint64)
        rate = np.ascontiguousarray(kind_rate, np.float64)
        n = raw.size
        out = np.zeros(n, np.float32)
        order = np.zeros(n, np.uint32) if rerank else None
        cfg = _capi.CxDecayConfig(int(enabled), max_age_days, min_factor, echo_weight, echo_cap)
        _check(self._L.cx_apply_score_decay(self._h, C.byref(cfg), C.c_float(recency_bias), n, int(seg_len),
                                            raw.ctypes.data, idle.ctypes.data, acc.ctypes.data, rate.ctypes.data,
                                            out.ctypes.data, order.ctypes.data if rerank else None))
        return (out, order) if rerank else out

    def row_id(self, row: int) -> bytes:
        buf = np.zeros(16, np.uint8)
        _check(self._L.cx_row_id(self._h, int(row), buf.ctypes.data))
        return buf.tobytes()

    # ---- persistence ----------------------------------------------------------
    def save(self, path: str) -> None:
        _check(self._L.cx_save(self._h, str(path).encode()))

    @classmethod
    def load(cls, path: str, device: int = 0, devices: Optional[Sequence[int]] = None) -> "GpuVectorIndex":
        L = _capi.load()
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            _check(L.cx_load_sharded(str(path).encode(), arr, len(devices), C.byref(h)))
        else:
            _check(L.cx_load(str(path).encode(), device, C.byref(h)))
        dim = int(L.cx_dimension(h))
        return cls(dim, device, _handle=h.value, devices=devices)

    # ---- instrumentation ------------------------------------------------------
    def stats(self) -> dict:
        s = _capi.CxStats()
        _check(self._L.cx_get_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in s._fields_}

    def set_option(self, key: str, value: int) -> None:
        _check(self._L.cx_set_option(self._h, key.encode(), int(value)))
