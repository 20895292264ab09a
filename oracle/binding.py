"""ctypes binding of the CPU oracle (oracle/cortex_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(cortex_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcortex_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("cortex_oracle.c", "hnsw_oracle.c", "aux_oracle.c", "cpu_fast.c", "Makefile")]
    fast = os.path.join(_HERE, "libcortex_cpufast_v3.so")
    stale = not os.path.exists(_LIB_PATH) or not os.path.exists(fast) or any(
        os.path.getmtime(s) > min(os.path.getmtime(_LIB_PATH), os.path.getmtime(fast)) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


class _Filter(C.Structure):
    _fields_ = [
        ("has_kinds", C.c_int),
        ("kinds", C.POINTER(C.c_char_p)),
        ("n_kinds", C.c_size_t),
        ("has_exclude", C.c_int),
        ("exclude", C.c_void_p),
        ("n_exclude", C.c_size_t),
        ("has_agent", C.c_int),
        ("agent", C.c_char_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    L.cxo_distance.restype = C.c_float
    L.cxo_distance.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.cxo_distance_to_similarity.restype = C.c_float
    L.cxo_distance_to_similarity.argtypes = [C.c_float]
    L.cxo_create.restype = C.c_void_p
    L.cxo_create.argtypes = [C.c_size_t]
    L.cxo_destroy.argtypes = [C.c_void_p]
    L.cxo_set_faithful_copy.argtypes = [C.c_void_p, C.c_int]
    L.cxo_insert.restype = C.c_int
    L.cxo_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.cxo_insert_batch.restype = C.c_int
    L.cxo_insert_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.cxo_remove.restype = C.c_int
    L.cxo_remove.argtypes = [C.c_void_p, C.c_void_p]
    L.cxo_set_metadata.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p]
    L.cxo_len.restype = C.c_size_t
    L.cxo_len.argtypes = [C.c_void_p]
    L.cxo_search.restype = C.c_size_t
    L.cxo_search.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cxo_search_threshold.restype = C.c_size_t
    L.cxo_search_threshold.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p,
                                       C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cxo_search_batch.restype = C.c_int
    L.cxo_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int]
    L.cxo_max_threads.restype = C.c_int
    L.cxo_save.restype = C.c_int
    L.cxo_save.argtypes = [C.c_void_p, C.c_char_p]
    L.cxo_load.restype = C.c_void_p
    L.cxo_load.argtypes = [C.c_char_p]
    _lib = L
    return L


def distance(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return float(lib().cxo_distance(a.ctypes.data, a.size, b.ctypes.data, b.size))


def distance_to_similarity(d: float) -> float:
    return float(lib().cxo_distance_to_similarity(C.c_float(d)))


@dataclass
class Filter:
    """Mirror of VectorFilter (vector/index.rs:17-47)."""
    kinds: Optional[Sequence[str]] = None
    exclude: Optional[Sequence[bytes]] = None
    source_agent: Optional[str] = None


@dataclass
class Hits:
    ids: np.ndarray       # [n,16] uint8
    score: np.ndarray     # [n] float32
    distance: np.ndarray  # [n] float32
    rows: np.ndarray      # [n] uint32 (oracle row = insertion slot)


def _id(b) -> np.ndarray:
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    assert a.size == 16
    return np.ascontiguousarray(a, dtype=np.uint8)


class OracleIndex:
    """Mirror of HnswIndex in brute-force mode (vector/index.rs:182-473)."""

    def __init__(self, dimension: int, faithful_copy: bool = True, _handle=None):
        self._L = lib()
        self._h = _handle if _handle is not None else self._L.cxo_create(dimension)
        self.dimension = dimension
        self._L.cxo_set_faithful_copy(self._h, int(faithful_copy))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.cxo_destroy(self._h)
            self._h = None

    def insert(self, node_id, embedding) -> None:
        v = np.ascontiguousarray(embedding, dtype=np.float32)
        i = _id(node_id)
        rc = self._L.cxo_insert(self._h, i.ctypes.data, v.ctypes.data, v.size)
        if rc != 0:
            raise ValueError(
                f"Embedding dimension mismatch: expected {self.dimension}, got {v.size}")

    def insert_batch(self, ids: np.ndarray, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.uint8).reshape(-1, 16)
        if rows.shape[0] == 0:
            return
        rc = self._L.cxo_insert_batch(self._h, ids.ctypes.data, rows.ctypes.data, rows.shape[0], rows.shape[1])
        if rc != 0:
            raise ValueError("Embedding dimension mismatch")

    def remove(self, node_id) -> None:
        self._L.cxo_remove(self._h, _id(node_id).ctypes.data)

    def set_metadata(self, node_id, kind: str, source_agent: str) -> None:
        self._L.cxo_set_metadata(self._h, _id(node_id).ctypes.data, kind.encode(), source_agent.encode())

    def __len__(self) -> int:
        return int(self._L.cxo_len(self._h))

    def _filter(self, f: Optional[Filter]):
        if f is None:
            return None, None
        keep = []
        cf = _Filter()
        if f.kinds is not None:
            arr = (C.c_char_p * max(1, len(f.kinds)))(*[k.encode() for k in f.kinds])
            keep.append(arr)
            cf.has_kinds, cf.kinds, cf.n_kinds = 1, C.cast(arr, C.POINTER(C.c_char_p)), len(f.kinds)
        if f.exclude is not None:
            ex = np.concatenate([_id(e) for e in f.exclude]) if len(f.exclude) else np.zeros(0, np.uint8)
            ex = np.ascontiguousarray(ex)
            keep.append(ex)
            cf.has_exclude, cf.exclude, cf.n_exclude = 1, ex.ctypes.data, len(f.exclude)
        if f.source_agent is not None:
            cf.has_agent, cf.agent = 1, f.source_agent.encode()
        keep.append(cf)
        return C.addressof(cf), keep

    def search(self, query, k: int, flt: Optional[Filter] = None) -> Hits:
        q = np.ascontiguousarray(query, dtype=np.float32)
        n_alloc = max(1, min(k, len(self)))
        ids = np.zeros((n_alloc, 16), np.uint8)
        sc = np.zeros(n_alloc, np.float32)
        di = np.zeros(n_alloc, np.float32)
        rw = np.zeros(n_alloc, np.uint32)
        fp, keep = self._filter(flt)
        n = self._L.cxo_search(self._h, q.ctypes.data, q.size, min(k, n_alloc), fp,
                               ids.ctypes.data, sc.ctypes.data, di.ctypes.data, rw.ctypes.data)
        del keep
        return Hits(ids[:n], sc[:n], di[:n], rw[:n])

    def search_threshold(self, query, threshold: float, flt: Optional[Filter] = None) -> Hits:
        q = np.ascontiguousarray(query, dtype=np.float32)
        cap = max(1, len(self))
        ids = np.zeros((cap, 16), np.uint8)
        sc = np.zeros(cap, np.float32)
        di = np.zeros(cap, np.float32)
        rw = np.zeros(cap, np.uint32)
        fp, keep = self._filter(flt)
        n = self._L.cxo_search_threshold(self._h, q.ctypes.data, q.size, C.c_float(threshold), fp, cap,
                                         ids.ctypes.data, sc.ctypes.data, di.ctypes.data, rw.ctypes.data)
        del keep
        return Hits(ids[:n].copy(), sc[:n].copy(), di[:n].copy(), rw[:n].copy())

    def search_batch(self, queries, k: int, flt: Optional[Filter] = None, n_threads: int = 0):
        """Returns (ids[B,k,16], score[B,k], distance[B,k], rows[B,k], n[B])."""
        Q = np.ascontiguousarray(queries, dtype=np.float32)
        B, qlen = Q.shape
        kk = max(1, min(k, max(1, len(self))))
        ids = np.zeros((B, kk, 16), np.uint8)
        sc = np.zeros((B, kk), np.float32)
        di = np.zeros((B, kk), np.float32)
        rw = np.zeros((B, kk), np.uint32)
        n = np.zeros(B, np.uint64)
        fp, keep = self._filter(flt)
        self._L.cxo_search_batch(self._h, Q.ctypes.data, B, qlen, kk, fp, ids.ctypes.data,
                                 sc.ctypes.data, di.ctypes.data, rw.ctypes.data, n.ctypes.data, n_threads)
        del keep
        return ids, sc, di, rw, n

    def rebuild(self) -> None:
        """vector/index.rs:416-435 builds the HNSW graph; the exact scan has nothing to build."""

    def save(self, path: str) -> None:
        if self._L.cxo_save(self._h, path.encode()) != 0:
            raise ValueError(f"Failed to write index file: {path}")

    @classmethod
    def load(cls, path: str) -> "OracleIndex":
        L = lib()
        h = L.cxo_load(path.encode())
        if not h:
            raise ValueError(f"Failed to read index file: {path}")
        L.cxo_dim.restype = C.c_size_t
        L.cxo_dim.argtypes = [C.c_void_p]
        return cls(int(L.cxo_dim(h)), _handle=h)


def walk_node(value: bytes, dim: int):
    """aux_oracle.c cxo_walk_node: (status, id bytes, embedding or None, created_ns, last_accessed_ns, access_count)."""
    L = lib()
    L.cxo_walk_node.restype = C.c_int
    L.cxo_walk_node.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64),
                                C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
    idb = np.zeros(16, np.uint8)
    row = np.zeros(max(1, dim), np.float32)
    cr, la, ac = C.c_int64(0), C.c_int64(0), C.c_uint64(0)
    st = L.cxo_walk_node(value, len(value), dim, idb.ctypes.data, row.ctypes.data, C.byref(cr), C.byref(la), C.byref(ac))
    return st, idb.tobytes(), (row[:dim].copy() if st == 0 else None), cr.value, la.value, ac.value


def apply_score_decay(raw, idle_seconds, access_count, kind_rate, enabled=True, max_age_days=365.0, min_factor=0.1,
                      echo_weight=0.05, echo_cap=2.0, recency_bias=0.15) -> float:
    """aux_oracle.c cxo_apply_score_decay (vector/scoring.rs:84-114); defaults = ScoreDecayConfig::default()."""
    L = lib()
    L.cxo_apply_score_decay.restype = C.c_float
    L.cxo_apply_score_decay.argtypes = [C.c_float, C.c_int64, C.c_uint64, C.c_double, C.c_int, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_float]
    return float(L.cxo_apply_score_decay(raw, int(idle_seconds), int(access_count), kind_rate, int(enabled), max_age_days,
                                         min_factor, echo_weight, echo_cap, recency_bias))


def max_threads() -> int:
    """Host threads the CPU legs use: every core this process may run on.  (Not omp_get_max_threads():
    torchrun exports OMP_NUM_THREADS=1, which would silently shrink the baseline to one core.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class OracleHnsw:
    """The HNSW branch of HnswIndex::search (index.rs:342-373) restated on the CPU
    (oracle/hnsw_oracle.c): PARITY UNPINNED, parameters unverified -- used only to report
    the reference's approximate path as recall@k against the exact scan."""

    def __init__(self, vectors: np.ndarray, M: int = 32, ef_construction: int = 100, ef_search: int = 100,
                 seed: int = 1, n_threads: int = 1):
        L = lib()
        L.cxo_hnsw_build_mt.restype = C.c_void_p
        L.cxo_hnsw_build_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                        C.c_int]
        L.cxo_hnsw_search.restype = C.c_size_t
        L.cxo_hnsw_search.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.cxo_hnsw_distance_evals.restype = C.c_uint64
        L.cxo_hnsw_distance_evals.argtypes = [C.c_void_p]
        L.cxo_hnsw_free.argtypes = [C.c_void_p]
        self._L = L
        self._v = np.ascontiguousarray(vectors, dtype=np.float32)  # borrowed by the C side
        self.M, self.ef_construction, self.ef_search = M, ef_construction, ef_search
        self._h = L.cxo_hnsw_build_mt(self._v.ctypes.data, self._v.shape[0], self._v.shape[1], M, ef_construction,
                                      ef_search, seed, n_threads)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.cxo_hnsw_free(self._h)
            self._h = None

    def search(self, query, k: int):
        """index.rs:345-371 without a filter: take(k*10) of the ascending results, first k."""
        q = np.ascontiguousarray(query, dtype=np.float32)
        m = min(self.ef_search, k * 10)
        ids = np.zeros(m, np.uint32)
        d = np.zeros(m, np.float32)
        n = self._L.cxo_hnsw_search(self._h, q.ctypes.data, m, ids.ctypes.data, d.ctypes.data)
        n = min(n, k)
        return ids[:n], d[:n]

    def distance_evals(self) -> int:
        return int(self._L.cxo_hnsw_distance_evals(self._h))


class CpuFastScan:
    """oracle/cpu_fast.c: an OPTIMISED CPU scan (hoisted norms, one corpus pass per 16 queries, vectorised
    reassociated fp32, k-heaps, OpenMP) -- a courtesy baseline, NOT the reference and not a checker."""

    def __init__(self, corpus: np.ndarray):
        build()
        flags = set()
        try:
            for line in open("/proc/cpuinfo"):
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    break
        except OSError:
            pass
        if {"avx512f", "avx512vl", "avx512bw", "avx512dq", "avx512cd"} <= flags:
            name = "libcortex_cpufast_v4.so"
        elif {"avx2", "fma", "bmi2"} <= flags:
            name = "libcortex_cpufast_v3.so"
        else:
            raise RuntimeError("host CPU has neither AVX-512 nor AVX2+FMA: no optimised CPU baseline")
        self.isa = name[len("libcortex_cpufast_"):-3]
        L = C.CDLL(os.path.join(_HERE, name))
        L.cxf_row_rnorms.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.cxf_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                       C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        self._L = L
        self._E = np.ascontiguousarray(corpus, dtype=np.float32)
        self._rn = np.zeros(self._E.shape[0], np.float32)
        L.cxf_row_rnorms(self._E.ctypes.data, self._E.shape[0], self._E.shape[1], self._rn.ctypes.data)

    def search_batch(self, queries: np.ndarray, k: int, n_threads: int = 0):
        Q = np.ascontiguousarray(queries, dtype=np.float32)
        rows = np.zeros((Q.shape[0], k), np.uint32)
        score = np.zeros((Q.shape[0], k), np.float32)
        self._L.cxf_search_batch(self._E.ctypes.data, self._rn.ctypes.data, self._E.shape[0], self._E.shape[1],
                                 Q.ctypes.data, Q.shape[0], k, rows.ctypes.data, score.ctypes.data, n_threads)
        return rows, score
