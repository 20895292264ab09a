#!/usr/bin/env python
"""Auto-link cycle breakdown: cx_autolink_batch_device for B new nodes x rows corpus, k=100.
usage: python scripts/al_probe.py [--rows N] [--new B] [--graphs 0|1] [--growth G]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cortex_b200 import GpuVectorIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--new", type=int, default=16384)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--graphs", default="1,0")
ap.add_argument("--growth", default="-1")
a = ap.parse_args()
dev = torch.device("cuda", 0)
corpus = bench.make_corpus_torch(a.rows, 384, bench.SEED, dev)
q = bench.make_queries_torch(corpus, a.new, bench.SEED + 5)
ix = GpuVectorIndex(384, device=0)
ix.insert_batch_device(bench.ids_for(a.rows), corpus)
del corpus
ix.set_option("profile", 1)
s = torch.cuda.current_stream().cuda_stream
for graphs in [int(x) for x in a.graphs.split(",")]:
    for growth in [int(x) for x in a.growth.split(",")]:
        ix.set_option("graphs", graphs)
        if growth >= 0:
            ix.set_option("tensor_phase_growth", growth)
        bufs = None
        for _ in range(3):
            res, bufs = ix.autolink_batch_device(q, a.k, 0.75, 50, stream=s, bufs=bufs)
        torch.cuda.synchronize()
        s0 = ix.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.reps):
            res, bufs = ix.autolink_batch_device(q, a.k, 0.75, 50, stream=s, bufs=bufs)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / a.reps * 1e3
        s1 = ix.stats()
        ms = e0.elapsed_time(e1) / a.reps
        scan_ms = (s1["pass_kernel_ns"] - s0["pass_kernel_ns"]) * 1e-6 / a.reps
        print(json.dumps({"graphs": graphs, "growth": growth, "ms_per_cycle": ms, "wall_ms": wall, "scan_ms": scan_ms,
                          "launches_per_cycle": (s1["kernel_launches"] - s0["kernel_launches"]) / a.reps,
                          "graph_launches": s1["graph_launches"] - s0["graph_launches"],
                          "fallbacks": s1["fallbacks"] - s0["fallbacks"],
                          "tflops_whole": 2.0 * 384 * a.new * a.rows / (ms * 1e-3) / 1e12}), flush=True)
