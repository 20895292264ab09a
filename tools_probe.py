import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from cortex_b200 import GpuVectorIndex, synth
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
corpus = bench.make_corpus_torch(n, 384, bench.SEED, dev)
q = bench.make_queries_torch(corpus, 256, bench.SEED)[:B].contiguous()
ids = np.zeros((n, 16), np.uint8); ids[:, 8:] = np.arange(n, dtype=np.uint64).astype(">u8").view(np.uint8).reshape(-1, 8)
ix = GpuVectorIndex(384); ix.insert_batch(ids, corpus.cpu().numpy())
qh = q.cpu().numpy()
for _ in range(5): ix.search_batch_arrays(qh, 10)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): ix.search_batch_arrays(qh, 10)
t1 = time.perf_counter()
print("host api us/step", (t1 - t0) / 20 * 1e6)
out = None
s = torch.cuda.current_stream().cuda_stream
for _ in range(5): out = ix.search_batch_device(q, 10, stream=s, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): out = ix.search_batch_device(q, 10, stream=s, out=out)
torch.cuda.synchronize()
t1 = time.perf_counter()
print("device api us/step", (t1 - t0) / 20 * 1e6)
