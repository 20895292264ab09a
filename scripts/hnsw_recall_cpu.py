#!/usr/bin/env python
"""CPU only: recall@k of the reference's approximate path (HNSW restated, oracle/hnsw_oracle.c -- parity
unpinned, parameters unverified) against the exact scan (oracle/cortex_oracle.c, which the GPU path
matches bit for bit).  usage: scripts/hnsw_recall_cpu.py [rows] [queries] > profiles/hnsw_recall_cpu.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle.binding import OracleHnsw, OracleIndex, max_threads  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 200
corpus = bench.make_corpus_torch(rows, 384, bench.SEED + 9, "cpu").numpy()
Q = bench.make_queries_torch(torch.from_numpy(corpus), nq, bench.SEED + 9).numpy()
ids = np.zeros((rows, 16), np.uint8)
ids[:, 8:] = np.arange(rows, dtype=np.uint64).astype(">u8").view(np.uint8).reshape(-1, 8)
ex = OracleIndex(384, faithful_copy=False)
ex.insert_batch(ids, corpus)
t0 = time.perf_counter()
hn = OracleHnsw(corpus)
build_s = time.perf_counter() - t0
out = {"workload": f"{rows} x 384 clustered unit-norm rows, {nq} queries", "host_threads": max_threads(),
       "hnsw": "restated: M=32, ef_construction=100, ef_search=100 (instant-distance 0.6.1 defaults from memory, "
               "unverified; parity unpinned)", "build_s_single_thread": build_s}
for k in (10, 100):
    _, _, _, erow, en = ex.search_batch(Q, k, n_threads=max_threads())
    t0 = time.perf_counter()
    got = [hn.search(q, k)[0] for q in Q]
    dt = time.perf_counter() - t0
    hit = sum(len(set(map(int, got[b])) & set(map(int, erow[b, :int(en[b])]))) for b in range(nq))
    out[f"recall@{k}"] = hit / float(en.sum())
    out[f"hnsw_queries_per_s_single_thread_k{k}"] = nq / dt
    out[f"results_per_query_k{k}"] = float(np.mean([len(x) for x in got]))
print(json.dumps(out, indent=1))
