#!/usr/bin/env python
"""The streaming pass alone at small batch sizes: kernel time per pass (CUDA events inside the library) and host
call time, bf16 shadow vs fp32 rows.  usage: python scripts/k1_probe.py [--batches 1,2,4] [--bf16 1,0]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cortex_b200 import GpuVectorIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--batches", default="1,2,4")
ap.add_argument("--bf16", default="1,0")
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
dev = torch.device("cuda", 0)
corpus = bench.make_corpus_torch(a.rows, a.dim, bench.SEED, dev)
q = bench.make_queries_torch(corpus, 8, bench.SEED)
ids = np.zeros((a.rows, 16), np.uint8)
ids[:, 8:] = np.arange(a.rows, dtype=np.uint64).astype(">u8").view(np.uint8).reshape(-1, 8)
ix = GpuVectorIndex(a.dim, device=0)
ix.reserve(a.rows)
ix.insert_batch_device(ids, corpus)
ix.set_option("profile", 1)
for half in [int(x) for x in a.bf16.split(",")]:
    ix.set_option("stream_bf16", half)
    for b in [int(x) for x in a.batches.split(",")]:
        qq = q[:b].contiguous()
        out = None
        for _ in range(3):
            out = ix.search_batch_device(qq, a.k, out=out)
        torch.cuda.synchronize()
        s0 = ix.stats()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            out = ix.search_batch_device(qq, a.k, out=out)
        torch.cuda.synchronize()
        call_us = (time.perf_counter() - t0) / a.reps * 1e6
        s1 = ix.stats()
        n = s1["pass_kernel_launches"] - s0["pass_kernel_launches"]
        us = (s1["pass_kernel_ns"] - s0["pass_kernel_ns"]) * 1e-3 / max(n, 1)
        bytes_ = a.rows * ((a.dim + 63) // 64 * 64) * 2 if half else a.rows * ((a.dim + 3) // 4 * 4) * 4
        print(json.dumps({"bf16": half, "B": b, "kernel_us": us, "gbs": bytes_ / us * 1e-3, "call_us": call_us,
                          "fallbacks": s1["fallbacks"] - s0["fallbacks"],
                          "bf16_queries": s1["queries_stream_bf16"] - s0["queries_stream_bf16"]}), flush=True)
