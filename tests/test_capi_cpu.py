"""CPU-side checks of the C ABI: the library loads without a GPU, exports every
symbol include/cortex_gpu.h declares, and refuses to compute without a device."""
import ctypes as C
import os
import re

import pytest

from cortex_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _capi.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cortex_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(cx_[a-z_0-9]+)\s*\(", src))


def test_every_header_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in cortex_gpu.h but not exported"
    assert syms == set(_capi.SYMBOLS), "ctypes table and header disagree"


def test_rust_ffi_crate_declares_exactly_the_header_symbols():
    """rust/cortex-gpu-sys cannot be compiled here (no cargo): at least keep its extern block in step with
    the header, symbol for symbol."""
    src = open(os.path.join(ROOT, "rust", "cortex-gpu-sys", "src", "lib.rs")).read()
    rust = set(re.findall(r"pub fn (cx_[a-z_0-9]+)\s*\(", src))
    assert rust == header_symbols(), (sorted(header_symbols() - rust), sorted(rust - header_symbols()))


def test_every_option_and_stats_field_is_documented_and_mirrored():
    """cx_set_option keys accepted by the library are all described in the header (the probe-only measurement
    hook is described as such), and the cx_stats fields of the header, the ctypes mirror and the Rust mirror
    are the same list in the same order (the struct is filled positionally across the FFI)."""
    hdr = open(os.path.join(ROOT, "include", "cortex_gpu.h")).read()
    src = open(os.path.join(ROOT, "cortex_b200", "csrc", "cx_index.cu")).read()
    keys = set(re.findall(r'strcmp\(key, "([a-z_0-9]+)"\)', src))
    assert len(keys) >= 10
    for k in keys:
        assert f'"{k}"' in hdr, f"option {k} is accepted by cx_set_option but not documented in cortex_gpu.h"
    body = re.search(r"typedef struct cx_stats \{(.*?)\} cx_stats;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r"uint64_t\s+([^;]+);", body) for f in decl.split(",")]
    assert [n for n, _ in _capi.CxStats._fields_] == fields
    rs = open(os.path.join(ROOT, "rust", "cortex-gpu-sys", "src", "lib.rs")).read()
    rs_body = re.search(r"pub struct cx_stats \{(.*?)\}", rs, flags=re.S).group(1)
    assert re.findall(r"pub ([a-z_0-9]+): u64", rs_body) == fields


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.cx_version()
    assert isinstance(lib.cx_last_error(), bytes)


def test_argument_validation_needs_no_device(lib):
    h = C.c_void_p()
    assert lib.cx_index_create(0, 0, C.byref(h)) == _capi.CX_ERR_VALIDATION
    assert b"dimension" in lib.cx_last_error()
    assert lib.cx_len(None) == 0


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    st = lib.cx_index_create(384, 0, C.byref(h))
    assert st == _capi.CX_ERR_CUDA
    assert b"no CPU fallback" in lib.cx_last_error()
    from cortex_b200 import CortexError, GpuVectorIndex
    with pytest.raises(CortexError):
        GpuVectorIndex(384)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; nothing under cortex_b200/ may reference it."""
    pkg = os.path.join(ROOT, "cortex_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "build.py" and False, f"{f} mentions the oracle"
