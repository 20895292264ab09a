#!/usr/bin/env python
"""How much do concurrent callers overlap?  T host threads each run device-resident searches (B=1024, k=10, 1M x 384)
on one index; prints ms per search for T = 1, 2, 3."""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cortex_b200 import GpuVectorIndex  # noqa: E402

dev = torch.device("cuda", 0)
corpus = bench.make_corpus_torch(1_000_000, 384, bench.SEED, dev)
q = bench.make_queries_torch(corpus, 1024, bench.SEED)
ix = GpuVectorIndex(384)
ix.insert_batch_device(bench.ids_for(1_000_000), corpus)
del corpus
for T in (1, 2, 3):
    n = 60
    streams = [torch.cuda.Stream() for _ in range(T)]
    start = threading.Barrier(T + 1)
    done = []

    def run(i):
        torch.cuda.set_device(0)
        out = None
        s = streams[i].cuda_stream
        for _ in range(5):
            out = ix.search_batch_device(q, 10, stream=s, out=out)
        start.wait()
        for _ in range(n):
            out = ix.search_batch_device(q, 10, stream=s, out=out)
        streams[i].synchronize()
        done.append(time.perf_counter())
    ths = [threading.Thread(target=run, args=(i,)) for i in range(T)]
    for t in ths:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in ths:
        t.join()
    print(json.dumps({"threads": T, "ms_per_search": (max(done) - t0) / (n * T) * 1e3}), flush=True)
