"""In-tree build of libcortex_gpu.so (CUDA kernels + C ABI) for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU
box with the gpurun snapshot.  cudart is linked statically so the library has no
CUDA runtime dependency at dlopen time (it loads on a CPU-only host; every entry
point that needs a device then fails with CX_ERR_CUDA).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libcortex_gpu.so")
STAMP = os.path.join(HERE, ".libcortex_gpu.stamp")
# measurement build for scripts/k2_probe.py only (-DCX_PROBE: the result-corrupting "tensor_debug" hook exists);
# never loaded by the package unless CORTEX_GPU_LIB points at it
PROBE_LIB = os.path.join(HERE, "libcortex_gpu_probe.so")
PROBE_STAMP = os.path.join(HERE, ".libcortex_gpu_probe.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-shared", "-cudart", "static",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cortex_b200 has no CPU fallback and cannot be built without it")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))
    ) + [os.path.join(INCLUDE, "cortex_gpu.h"), os.path.abspath(__file__)]
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fp:
            h.update(fp.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, probe: bool = False) -> str:
    dig = _digest()
    lib, stamp = (PROBE_LIB, PROBE_STAMP) if probe else (LIB, STAMP)
    if not force and os.path.exists(lib) and os.path.exists(stamp):
        if open(stamp).read().strip() == dig:
            return lib
    cmd = [_nvcc(), *NVCC_FLAGS, *(["-DCX_PROBE"] if probe else []), "-I", INCLUDE, "-I", CSRC, "-o", lib, *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without libgomp specs; use the system compiler
    cmd[1:1] = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose:
        print(r.stderr, file=sys.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fp:
        fp.write(dig)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, probe="--probe" in sys.argv))
