"""ctypes declarations for include/cortex_gpu.h (the C ABI).

This is the Python counterpart of the Rust `cortex-gpu-sys` crate sketched in
INTEGRATION.md: it declares exactly the symbols the header exports and nothing
else.  Loading fails loudly if the library is missing or a symbol is absent;
there is no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CORTEX_GPU_LIB: load another build of the same sources (the -DCX_PROBE measurement build of scripts/k2_probe.py)
LIB_PATH = os.environ.get("CORTEX_GPU_LIB") or os.path.join(_HERE, "libcortex_gpu.so")

CX_OK, CX_ERR_VALIDATION, CX_ERR_CUDA, CX_ERR_NCCL, CX_ERR_IO = 0, 1, 2, 3, 4


class CxFilter(C.Structure):
    _fields_ = [
        ("has_kinds", C.c_int32),
        ("kinds", C.POINTER(C.c_char_p)),
        ("n_kinds", C.c_uint32),
        ("has_exclude", C.c_int32),
        ("exclude_ids", C.c_void_p),
        ("n_exclude", C.c_uint32),
        ("has_source_agent", C.c_int32),
        ("source_agent", C.c_char_p),
    ]


class CxDecayConfig(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("max_age_days", C.c_double), ("min_factor", C.c_double),
                ("echo_weight", C.c_double), ("echo_cap", C.c_double)]


class CxStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64),
        ("queries_stream", C.c_uint64),
        ("queries_tensor", C.c_uint64),
        ("queries_exact", C.c_uint64),
        ("fallbacks", C.c_uint64),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("pass_kernel_ns", C.c_uint64),
        ("pass_kernel_launches", C.c_uint64),
        ("graph_launches", C.c_uint64),
        ("grow_events", C.c_uint64),
        ("irregular_rows", C.c_uint64),
        ("capacity_rows", C.c_uint64),
        ("in_place_growth", C.c_uint64),
        ("grow_ns", C.c_uint64),
        ("grow_ns_max", C.c_uint64),
        ("grow_waits", C.c_uint64),
        ("unverified_overflow", C.c_uint64),
        ("unverified_near_ties", C.c_uint64),
        ("unverified_other", C.c_uint64),
        ("queries_stream_bf16", C.c_uint64),
    ]


# name -> (restype, argtypes); kept in one table so tests can check that every
# symbol the header declares is exported.
SYMBOLS = {
    "cx_index_create": (C.c_int, [C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]),
    "cx_index_create_sharded": (C.c_int, [C.c_uint32, C.POINTER(C.c_int), C.c_uint32, C.POINTER(C.c_void_p)]),
    "cx_shard_count": (C.c_uint32, [C.c_void_p]),
    "cx_index_destroy": (None, [C.c_void_p]),
    "cx_insert": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "cx_insert_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]),
    "cx_insert_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]),
    "cx_remove": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cx_set_metadata": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p]),
    "cx_len": (C.c_uint64, [C.c_void_p]),
    "cx_dimension": (C.c_uint32, [C.c_void_p]),
    "cx_rebuild": (C.c_int, [C.c_void_p]),
    "cx_reserve": (C.c_int, [C.c_void_p, C.c_uint64]),
    "cx_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "cx_search_threshold": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_void_p, C.c_uint64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64)]),
    "cx_search_threshold_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_float, C.c_void_p,
                                            C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cx_dedup_scan": (C.c_int, [C.c_void_p, C.c_float, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "cx_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cx_search_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "cx_search_batch_device_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.POINTER(C.c_void_p)]),
    "cx_search_batch_device_end": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "cx_autolink_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64,
                                    C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cx_autolink_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_float, C.c_uint32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cx_pack_topk_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "cx_merge_topk_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cx_autolink_filter_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                            C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cx_search_ticket_ok": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "cx_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cx_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "cx_load_sharded": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_uint32, C.POINTER(C.c_void_p)]),
    "cx_row_id": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "cx_extract_embeddings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cx_load_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "cx_apply_score_decay": (C.c_int, [C.c_void_p, C.POINTER(CxDecayConfig), C.c_float, C.c_uint64, C.c_uint32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cx_get_stats": (C.c_int, [C.c_void_p, C.POINTER(CxStats)]),
    "cx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "cx_debug_tensor_plan": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p,
                                       C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_double)]),
    "cx_last_error": (C.c_char_p, []),
    "cx_version": (C.c_char_p, []),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -m cortex_b200.build` (needs nvcc). "
            "cortex_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
