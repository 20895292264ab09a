#!/usr/bin/env python
"""Host-to-host latency of search_batch at small batch sizes, pageable vs pinned query buffers."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cortex_b200 import GpuVectorIndex
rows = 1_000_000
dev = torch.device("cuda", 0)
corpus = bench.make_corpus_torch(rows, 384, bench.SEED, dev)
q_all = bench.make_queries_torch(corpus, 1024, bench.SEED)
ids = np.zeros((rows, 16), np.uint8)
ids[:, 8:] = np.arange(rows, dtype=np.uint64).astype(">u8").view(np.uint8).reshape(-1, 8)
ix = GpuVectorIndex(384); ix.reserve(rows); ix.insert_batch_device(ids, corpus)
graphs = int(os.environ.get("CX_GRAPHS", "1"))
ix.set_option("graphs", graphs)
if "CX_TMB" in os.environ:
    ix.set_option("tensor_min_batch", int(os.environ["CX_TMB"]))
for B in [int(x) for x in os.environ.get("CX_BATCHES", "1,2,4,8,64,1024").split(",")]:
    hq = q_all[:B].cpu().numpy()
    hp = torch.empty((B, 384), dtype=torch.float32).pin_memory(); hp.copy_(q_all[:B]); hpn = hp.numpy()
    for name, buf in (("pageable", hq), ("pinned", hpn)):
        for _ in range(3): ix.search_batch_arrays(buf, 10)
        t0 = time.perf_counter(); n = 200 if B <= 64 else 50
        for _ in range(n): ix.search_batch_arrays(buf, 10)
        dt = (time.perf_counter() - t0) / n
        print(json.dumps({"graphs": graphs, "B": B, "buf": name, "ms": dt * 1e3, "qps": B / dt}), flush=True)
