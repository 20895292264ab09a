#!/usr/bin/env python
"""configs[3] shard shape (1024-d bf16-valued rows, B=256, top-100): which path answers, how many queries need a
retry, what a step costs.  usage: python scripts/cfg4_probe.py [rows]"""
import json
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from cortex_b200 import GpuVectorIndex  # noqa: E402

dev = torch.device("cuda", 0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
d, B, k = 1024, 256, 100
ix = GpuVectorIndex(d)
ix.reserve(rows)
q = None
for s0 in range(0, rows, 625_000):
    n = min(625_000, rows - s0)
    c = bench.make_corpus_torch(n, d, bench.SEED + 17 * (s0 // 625_000), dev, cluster_seed=bench.SEED + 4,
                                n_clusters=50_000_000 // 1024)
    c = c.to(torch.bfloat16).to(torch.float32)
    ix.insert_batch_device(bench.ids_for(n, s0), c)
    if q is None:
        q = bench.make_queries_torch(c, B, bench.SEED)
    del c
s = torch.cuda.current_stream().cuda_stream
ix.set_option("profile", 1)
import os
pairs = [int(x) for x in os.environ.get("CX_PAIRS", "0").split(",")]
epis = [int(x) for x in os.environ.get("CX_EPI", "8").split(",")]
for force, pair, epi in [(0, p_, e_) for p_ in pairs for e_ in epis] + ([(1, 0, 8)] if os.environ.get("CX_FORCE_STREAM") else []):
    ix.set_option("force_path", force)
    ix.set_option("tensor_pair", pair)
    ix.set_option("tensor_epi_warps", epi)
    out = None
    for _ in range(2):
        out = ix.search_batch_device(q, k, stream=s, out=out)
    torch.cuda.synchronize()
    st0 = ix.stats()
    t0 = time.perf_counter()
    reps = 6
    for _ in range(reps):
        out = ix.search_batch_device(q, k, stream=s, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps * 1e3
    st1 = ix.stats()
    scan_us = (st1["pass_kernel_ns"] - st0["pass_kernel_ns"]) * 1e-3 / max(1, st1["pass_kernel_launches"] - st0["pass_kernel_launches"])
    print(json.dumps({"force_path": force, "pair": pair, "epi": epi, "ms": dt, "scan_us": scan_us,
                      "tflops": 2.0 * d * B * rows / (scan_us * 1e-6) / 1e12,
                      **{kk: (st1[kk] - st0[kk]) / reps for kk in
                      ("fallbacks", "queries_stream", "queries_tensor", "queries_exact", "kernel_launches")},
                      "why": [st1[w] - st0[w] for w in ("unverified_overflow", "unverified_near_ties", "unverified_other")]}), flush=True)
