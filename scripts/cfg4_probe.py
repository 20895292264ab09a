import sys, time, json, torch
sys.path.insert(0, "/root/repo")
import bench
from cortex_b200 import GpuVectorIndex
dev = torch.device("cuda", 0)
rows, d, B, k = 1_000_000, 1024, 256, 100
ix = GpuVectorIndex(d)
q = None
for s0 in range(0, rows, 500_000):
    c = bench.make_corpus_torch(500_000, d, 5 + s0, dev).to(torch.bfloat16).to(torch.float32)
    ix.insert_batch_device(bench.ids_for(500_000, s0), c)
    if q is None:
        q = bench.make_queries_torch(c, B, 3)
    del c
s = torch.cuda.current_stream().cuda_stream
ix.set_option("profile", 1)
for graphs in (1, 0):
    ix.set_option("graphs", graphs)
    out = None
    for _ in range(3):
        out = ix.search_batch_device(q, k, stream=s, out=out)
    torch.cuda.synchronize()
    st0 = ix.stats(); t0 = time.perf_counter()
    for _ in range(10):
        out = ix.search_batch_device(q, k, stream=s, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10 * 1e3
    st1 = ix.stats()
    print(json.dumps({"mode": "one-call", "graphs": graphs, "ms": dt, "scan_ms": (st1["pass_kernel_ns"] - st0["pass_kernel_ns"]) / 10e6,
                      "fallbacks": st1["fallbacks"] - st0["fallbacks"], "stream": st1["queries_stream"] - st0["queries_stream"], "exact": st1["queries_exact"] - st0["queries_exact"], "launches": (st1["kernel_launches"] - st0["kernel_launches"]) / 10}), flush=True)
    outs = [None, None]
    pend = []
    torch.cuda.synchronize(); st0 = ix.stats(); t0 = time.perf_counter()
    for i in range(10):
        outs[i % 2], t = ix.search_batch_device_begin(q, k, stream=s, out=outs[i % 2])
        pend.append(t)
        if len(pend) >= 2:
            ix.search_batch_device_end(pend.pop(0))
    while pend:
        ix.search_batch_device_end(pend.pop(0))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10 * 1e3
    st1 = ix.stats()
    print(json.dumps({"mode": "begin/end", "graphs": graphs, "ms": dt, "scan_ms": (st1["pass_kernel_ns"] - st0["pass_kernel_ns"]) / 10e6,
                      "fallbacks": st1["fallbacks"] - st0["fallbacks"], "stream": st1["queries_stream"] - st0["queries_stream"]}), flush=True)
