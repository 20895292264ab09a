#!/usr/bin/env python
"""Where does the tensor pass spend its time?  Times tensor_scan_kernel alone (CUDA events
inside the library) with the epilogue progressively disabled, CTA pair vs single CTA.
usage: python scripts/k2_probe.py [--rows N] [--batch B] [--k K]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the measurement hook "tensor_debug" exists only in the -DCX_PROBE build (python -m cortex_b200.build --probe)
os.environ.setdefault("CORTEX_GPU_LIB", os.path.join(ROOT, "cortex_b200", "libcortex_gpu_probe.so"))
import bench  # noqa: E402
from cortex_b200 import GpuVectorIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--debug-modes", default="0,2,1")
ap.add_argument("--growth", default="8")
ap.add_argument("--pairs", default="1,0")
ap.add_argument("--epi", default="8")
ap.add_argument("--sample", default="0")
ap.add_argument("--leftover", default="1")
a = ap.parse_args()
dev = torch.device("cuda", 0)
corpus = bench.make_corpus_torch(a.rows, 384, bench.SEED, dev)
q = bench.make_queries_torch(corpus, a.batch, bench.SEED)
ids = np.zeros((a.rows, 16), np.uint8)
ids[:, 8:] = np.arange(a.rows, dtype=np.uint64).astype(">u8").view(np.uint8).reshape(-1, 8)
ix = GpuVectorIndex(384, device=0)
ix.reserve(a.rows)
ix.insert_batch_device(ids, corpus)
ix.set_option("profile", 1)
ix.set_option("force_path", 2)
out = None
for pair, growth, epi, sample, leftover in [(int(p), int(g), int(e), int(sm), int(lo)) for p in a.pairs.split(",")
                                            for g in a.growth.split(",") for e in a.epi.split(",")
                                            for sm in a.sample.split(",") for lo in a.leftover.split(",")]:
    for dbg in [int(x) for x in a.debug_modes.split(",")]:
        ix.set_option("tensor_pair", pair)
        if growth >= 0:
            ix.set_option("tensor_phase_growth", growth)   # < 0: leave the automatic choice
        ix.set_option("tensor_epi_warps", epi)
        ix.set_option("tensor_sample_tiles", sample)
        ix.set_option("tensor_leftover_sms", leftover)
        if dbg or a.debug_modes != "0":
            ix.set_option("tensor_debug", dbg)   # exists in the -DCX_PROBE build only
        reps = a.reps if dbg == 0 else 2
        out = ix.search_batch_device(q, a.k, out=out)
        torch.cuda.synchronize()
        s0 = ix.stats()
        import time
        t0 = time.perf_counter()
        for _ in range(reps):
            out = ix.search_batch_device(q, a.k, out=out)
        torch.cuda.synchronize()
        call_us = (time.perf_counter() - t0) / reps * 1e6
        s1 = ix.stats()
        ns = s1["pass_kernel_ns"] - s0["pass_kernel_ns"]
        n = s1["pass_kernel_launches"] - s0["pass_kernel_launches"]
        us = ns * 1e-3 / max(n, 1)
        tf = 2.0 * 384 * a.batch * a.rows / (us * 1e-6) / 1e12
        if dbg:
            ix.set_option("tensor_debug", -1)  # prints the effective SM clock of the last launch to stderr
        print(json.dumps({"call_us": call_us, "sample": sample, "leftover": leftover, "epi": epi, "pair": pair, "growth": growth, "debug": dbg, "fallbacks": s1["fallbacks"] - s0["fallbacks"], "us_per_launch": us, "tflops": tf, "launches": n,
                          "batch": a.batch, "k": a.k}), flush=True)
if a.debug_modes != "0":
    ix.set_option("tensor_debug", 0)
ix.set_option("tensor_pair", 0)
