#!/usr/bin/env python
"""bench.py -- the similarity-scan hot path on N B200s (driver contract in DESIGN.md §5).

Workload of the headline line (BASELINE.json configs[1]): search over a 1,000,000 x 384 fp32 corpus
per GPU, query batch B = 1024, top-10.  A "step" is one search_batch of B queries.

  value    : device-resident searches (cx_search_batch_device_begin/_end, two in flight), CUDA events,
             max over ranks.  N > 1 is WEAK scaling -- one 1M-row shard per GPU, one process per GPU,
             local top-k lists exchanged with one NCCL all_gather and merged -- so `value` counts
             per-shard query scans (B x N per step); the true queries/s against the N x 1M-row corpus
             (B / step) and the scored pairs/s are reported next to it.
  e2e      : the same metric through the reference-facing host call (cx_search_batch: host query
             buffer in, host ids / scores out, copies inside the timed region)
  parity   : the results of the timed calls against the CPU oracle on the SAME corpus and queries
             (ids equal, score / distance bits equal), at every N
  roofline : the scan pass (bootstrap included), algorithmic flops or bytes over its CUDA-event
             time measured inside the timed region, against MEASURED_PEAKS.json
  strong_scaling : BASELINE.json configs[2] as named -- auto-link cycle, 100k new nodes x 10M-row
             corpus split over the N GPUs, k = 100, threshold 0.75 -- scored pairs/s at this N
  cfg4     : (N = 8) configs[3] as named: 50M x 1024-d bf16-valued corpus over 8 GPUs, B = 256, top-100
  single_process : (N > 1) the same search through ONE process driving all N GPUs behind the C ABI
             (cx_index_create_sharded), measured by rank 0 while the other ranks idle
  cpu_baseline / --impl reference : the CPU restatement of the reference's exact scan (oracle/) on
             all host cores, plus the reference's HNSW path (restated, parity unpinned) with recall@k
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xC027E5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000, help="corpus rows per GPU")
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--no-small-probe", action="store_true", help="skip the extra B=1 measurement")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-queries", type=int, default=0, help="queries per CPU-baseline step (0 = one per thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hnsw", action="store_true", help="skip the CPU HNSW recall baseline")
    ap.add_argument("--hnsw-rows", type=int, default=1_000_000)
    ap.add_argument("--autolink-new", type=int, default=16384,
                    help="new nodes per auto-link cycle in the extra 'autolink' measurement (0 = skip)")
    ap.add_argument("--strong-rows", type=int, default=10_000_000,
                    help="total corpus rows of the strong-scaling auto-link cycle, split over the GPUs (0 = skip)")
    ap.add_argument("--strong-new", type=int, default=100_000, help="new nodes of the strong-scaling auto-link cycle")
    ap.add_argument("--cfg4", default="auto", choices=["auto", "on", "off"],
                    help="configs[3]: 6.25M x 1024-d rows per GPU, B=256, top-100 (auto = at 8 GPUs)")
    ap.add_argument("--no-single-process", action="store_true")
    ap.add_argument("--pipeline-depth", type=int, default=2,
                    help="searches in flight in the device-resident measurement (1 = strictly one after the other)")
    ap.add_argument("--pin", action="store_true",
                    help="N > 1: give every rank its own contiguous block of the host's CPUs (sched_setaffinity)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=INT",
                    help="cx_set_option before the run (A/B measurements), e.g. --opt graphs=0")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j["bf16_tflops"]),
                "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(kernel, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu
    --set full capture of this same command (profiles/ncu_traffic.json, written by
    scripts/ncu_summary.py --traffic); None if there is no capture for this kernel / batch."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        j = json.load(open(p))
        e = j.get(kernel)
        if e and int(e.get("batch", batch)) == int(batch):
            return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def workload_config(a, world):
    """The `config` object; identical for both arms so that the driver can tell they ran the same thing."""
    return {"workload": (f"search: {a.rows}x{a.dim}-d fp32 corpus per GPU, query batch {a.batch}, top-{a.k} "
                         f"(BASELINE.json configs[1])"),
            "rows_per_gpu": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k, "seed": SEED,
            "parallelism": f"row-shard x{world}",
            "cache": "inputs larger than L2: the 1.5 GB corpus shard is streamed from HBM every step",
            "value_counts": "per-shard query scans (batch x n_gpus per step); equals queries/s at 1 GPU"}


def ids_for(n, start=0):
    ids = np.zeros((n, 16), np.uint8)
    ids[:, 8:] = (np.arange(n, dtype=np.uint64) + start).astype(">u8").view(np.uint8).reshape(-1, 8)
    return ids


# ----------------------------------------------------------------------------------
# synthetic data (clustered unit-norm rows), generated with torch so
# that 1M x 384 takes a fraction of a second on the GPU and a few seconds on CPU.
def make_corpus_torch(n, d, seed, device, cluster_seed=None, n_clusters=None):
    """n clustered unit-norm rows.  cluster_seed / n_clusters: draw the centroids from their own seed, so that
    several calls (the shards of one corpus, the chunks of a big one) sample the SAME clusters with different
    noise -- a row-partitioned corpus, not unrelated corpora side by side."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n_clusters = max(1, n // 64) if n_clusters is None else max(1, int(n_clusters))
    gc = g
    if cluster_seed is not None:
        gc = torch.Generator(device=device)
        gc.manual_seed(cluster_seed)
    cent = torch.randn((n_clusters, d), generator=gc, device=device, dtype=torch.float32)
    cent /= cent.norm(dim=1, keepdim=True)
    out = torch.empty((n, d), device=device, dtype=torch.float32)
    chunk = 1 << 18
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        which = torch.randint(0, n_clusters, (e - s,), generator=g, device=device)
        # noise per coordinate scaled so that the intra-cluster cosines (0.7 .. 0.99) do not depend on the dimension
        scale = 0.035 * (384.0 / d) ** 0.5 * (0.3 + 1.3 * torch.rand((e - s, 1), generator=g, device=device))
        x = torch.randn((e - s, d), generator=g, device=device, dtype=torch.float32) * scale + cent[which]
        x /= x.norm(dim=1, keepdim=True)
        out[s:e] = x
    n_dup = n // 100  # 1 % exact duplicates of earlier rows
    if n_dup and n > 2:
        dst = torch.randint(1, n, (n_dup,), generator=g, device=device)
        src = (dst.double() * torch.rand((n_dup,), generator=g, device=device, dtype=torch.float64)).long()
        out[dst] = out[src]
    return out


def make_queries_torch(corpus, b, seed):
    import torch

    g = torch.Generator(device=corpus.device)
    g.manual_seed(seed + 1)
    pick = torch.randint(0, corpus.shape[0], (b,), generator=g, device=corpus.device)
    q = corpus[pick].clone()
    s = torch.where(torch.arange(b, device=corpus.device) % 2 == 0, 0.02, 0.06)[:, None]
    q += torch.randn(q.shape, generator=g, device=corpus.device) * s
    q /= q.norm(dim=1, keepdim=True)
    return q.contiguous()


# ----------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 100 ms during the timed region, for the first `n_gpus` GPUs of the box."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, n_gpus=1):
        self.n_gpus = n_gpus
        self.rows = []
        self.proc = None

    def start(self):
        try:
            ids = ",".join(str(i) for i in range(self.n_gpus))
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={ids}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        per = {}
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                g = per.setdefault(int(f[0]), {"sm": [], "mx": [], "w": [], "reasons": set()})
                g["sm"].append(float(f[1]))
                g["mx"].append(float(f[2]))
                g["w"].append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    g["reasons"].add(nm)
        if not per:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        gpus = [{"gpu": i, "sm_mhz": float(np.median(g["sm"])), "power_w": float(np.median(g["w"])),
                 "reasons": sorted(g["reasons"])} for i, g in sorted(per.items())]
        all_sm = [x["sm_mhz"] for x in gpus]
        out = {"sm_mhz": float(np.median(all_sm)), "sm_max_mhz": max(max(g["mx"]) for g in per.values()),
               "samples": len(per[min(per)]["sm"]), "reasons": sorted(set().union(*[g["reasons"] for g in per.values()]))}
        if len(gpus) > 1:
            out["sm_mhz_min_over_gpus"] = min(all_sm)
            out["per_gpu"] = gpus
        return out


# ----------------------------------------------------------------------------------
# CPU legs (rank 0 only).  Everything here is the CHECKER or the BASELINE, never the product path.
def cpu_reference_leg(ix, queries_np, k, steps, warmup, n_queries):
    """The reference's exact scan restated on the CPU (oracle/), all host threads, one bounded batch per
    step.  Returns (queries/s, cores, ms_per_step, sample)."""
    from oracle.binding import max_threads

    cores = max_threads()
    nq = min(n_queries or cores, queries_np.shape[0])
    times = []
    for s in range(warmup + steps):
        q = queries_np[(np.arange(nq) + s * nq) % queries_np.shape[0]]
        t0 = time.perf_counter()
        ix.search_batch(q, k, n_threads=cores)
        t1 = time.perf_counter()
        if s >= warmup:
            times.append(t1 - t0)
    tot = sum(times)
    qps = nq * len(times) / tot
    sample = (f"{nq} queries/step x {len(times)} steps against the full {len(ix)}-row corpus on {cores} host threads, "
              f"OpenMP over queries (= rayon par_iter, index.rs:397), per-pair row clone + full stable sort")
    return qps, cores, 1e3 * tot / len(times), sample


def cpu_optimised_leg(corpus_np, queries_np, k, n_queries=64):
    """Courtesy number, NOT the reference: oracle/cpu_fast.c (hoisted norms, one corpus pass per 16
    queries, AVX-512 reassociated fp32, k-heaps, OpenMP) on the same corpus.  Never raises."""
    try:
        from oracle.binding import CpuFastScan, max_threads

        f = CpuFastScan(corpus_np)
        q = queries_np[:n_queries]
        f.search_batch(q[:16], k, n_threads=max_threads())
        t0 = time.perf_counter()
        f.search_batch(q, k, n_threads=max_threads())
        dt = time.perf_counter() - t0
        return {"value": q.shape[0] / dt, "unit": "queries/s", "cores": max_threads(), "isa": "x86-64-" + f.isa,
                "sample": f"{q.shape[0]} queries against the full corpus, top-{k}",
                "note": "optimised CPU scan (oracle/cpu_fast.c), approximate scores, NOT the reference's algorithm"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:200]}


def cpu_hnsw_leg(corpus_np, queries_np, exact_rows_10, exact_rows_100, rows):
    """The reference's approximate path (index.rs:342-373: instant-distance HnswMap, restated in
    oracle/hnsw_oracle.c -- PARITY UNPINNED, parameters unverified) on the bench corpus: build time, q/s and
    recall@10 / @100 against the exact result.  Never raises."""
    try:
        from oracle.binding import OracleHnsw, max_threads

        n = min(rows, corpus_np.shape[0])
        # bounded: a 50k-row build predicts the full one (cost ~ n log n); if that would not fit the budget the
        # index is built over the largest prefix that does, and the line says so
        budget_s, probe = 110.0, min(50_000, n)
        t0 = time.perf_counter()
        hn = OracleHnsw(corpus_np[:probe], n_threads=max_threads())
        t_probe = time.perf_counter() - t0
        if n > probe:
            per_row = t_probe / probe
            while n > probe and per_row * n * (1.0 + 0.12 * np.log2(n / probe)) > budget_s:
                n //= 2
            del hn
            t0 = time.perf_counter()
            hn = OracleHnsw(corpus_np[:n], n_threads=max_threads())
        build_s = time.perf_counter() - t0
        nq = min(64, queries_np.shape[0], exact_rows_10.shape[0])
        hit10 = hit100 = 0
        t0 = time.perf_counter()
        got = [hn.search(queries_np[b], 100)[0] for b in range(nq)]
        dt = time.perf_counter() - t0
        for b in range(nq):
            hit10 += len(set(got[b][:10].tolist()) & set(exact_rows_10[b].tolist()))
            hit100 += len(set(got[b][:100].tolist()) & set(exact_rows_100[b].tolist()))
        if n < corpus_np.shape[0]:  # recall is measured against the exact top-k of the SAME prefix
            from oracle.binding import CpuFastScan

            fs = CpuFastScan(corpus_np[:n])
            ex, _ = fs.search_batch(queries_np[:nq], 100, n_threads=max_threads())
            exact_rows_10, exact_rows_100 = ex[:, :10], ex
            hit10 = hit100 = 0
            for b in range(nq):
                hit10 += len(set(got[b][:10].tolist()) & set(exact_rows_10[b].tolist()))
                hit100 += len(set(got[b][:100].tolist()) & set(exact_rows_100[b].tolist()))
        return {"rows": n, "rows_requested": rows, "build_s": build_s, "build_threads": max_threads(), "queries_per_s_single_thread": nq / dt,
                "recall_at_10": hit10 / (10.0 * nq), "recall_at_100": hit100 / (100.0 * nq), "queries": nq,
                "params": {"M": hn.M, "ef_construction": hn.ef_construction, "ef_search": hn.ef_search},
                "note": "restated instant-distance 0.6 HNSW, parity unpinned, parameters unverified; exact result = the "
                        "GPU path's (oracle-checked) top-k"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:300]}


def compare_with_oracle(o_res, g_rows, g_score, g_dist, g_n):
    """ids (as global rows) equal, score / distance bits equal, for every checked query."""
    oi, osc, od, orow, on = o_res
    nq = orow.shape[0]
    ok = bool(np.array_equal(np.asarray(g_n[:nq], np.int64), np.asarray(on, np.int64)))
    for b in range(nq):
        n = int(on[b])
        ok = ok and bool(np.array_equal(np.asarray(g_rows[b, :n], np.int64), np.asarray(orow[b, :n], np.int64)))
        for g_, o_ in ((g_score, osc), (g_dist, od)):
            a, c = np.asarray(g_[b, :n], np.float32), np.asarray(o_[b, :n], np.float32)
            ok = ok and bool(np.all((a.view(np.uint32) == c.view(np.uint32)) | (np.isnan(a) & np.isnan(c))))
    return ok


def rows_from_ids(ids):
    """bench ids are the global row number, big endian, in the low 8 bytes"""
    return np.ascontiguousarray(ids[..., 8:]).view(">u8").reshape(ids.shape[:-1]).astype(np.int64)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library
    chatter) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def reference_arm(a, world):
    """--impl reference: the reference's own CPU implementation of the path (restated: the reference is
    Rust and cannot be built here) on all host cores, same config / metric / unit as our arm."""
    import torch  # noqa: F401

    from oracle.binding import OracleIndex

    corpus = make_corpus_torch(a.rows, a.dim, SEED, "cpu")
    queries = make_queries_torch(corpus, max(256, a.batch), SEED).numpy()
    ix = OracleIndex(a.dim, faithful_copy=True)
    ix.insert_batch(ids_for(a.rows), corpus.numpy())
    qps, cores, ms, sample = cpu_reference_leg(ix, queries, a.k, a.steps, a.warmup, a.cpu_queries)
    if world > 1:
        sample += (f"; at {world} GPUs the corpus is {world} such shards: a query costs {world} shard scans, so "
                   f"shard scans/s (the unit of `value`) is unchanged and queries/s is value / {world}")
    emit({
        "impl": "reference", "metric": "queries/s", "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
        "queries_per_s": qps / world,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the run
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank == 0:
            reference_arm(a, world)
        return

    if a.pin and world > 1 and hasattr(os, "sched_setaffinity"):
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // world)
        os.sched_setaffinity(0, set(cpus[local_rank * per:(local_rank + 1) * per]))

    import datetime

    import torch
    import torch.distributed as dist

    from cortex_b200 import GpuVectorIndex
    from cortex_b200.sharded import ShardedSearch

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # host-side barrier for the phases in which only rank 0 works (an NCCL barrier would spin on every GPU)
        host_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=30))

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.current_stream()
    pk = peaks()

    # ---- data: this rank's shard + the (shared) query batch ---------------------
    # one clustered corpus, row-partitioned: every rank draws from the same clusters (its own noise), so a
    # query has near neighbours in every shard, as it would in a corpus dealt out row by row
    def shard_corpus(r):
        return make_corpus_torch(a.rows, a.dim, SEED + 7919 * r, dev, cluster_seed=SEED, n_clusters=a.rows // 64)

    corpus = shard_corpus(rank)
    q_all = make_queries_torch(corpus, max(256, a.batch), SEED)
    if world > 1:
        dist.broadcast(q_all, src=0)
    d_q = q_all[:a.batch].contiguous()
    q_al = make_queries_torch(corpus, a.autolink_new, SEED + 5) if a.autolink_new > 0 else None
    ix = GpuVectorIndex(a.dim, device=local_rank)
    ix.insert_batch_device(ids_for(a.rows, rank * a.rows), corpus)
    corpus_np = corpus.cpu().numpy() if rank == 0 else None
    del corpus
    torch.cuda.empty_cache()
    ix.set_option("profile", 1)
    for kv in a.opt:
        key, val = kv.split("=")
        ix.set_option(key, int(val))

    out = None

    def local_search(q, k):
        nonlocal out
        out = ix.search_batch_device(q, k, stream=stream.cuda_stream, out=out)
        return out

    outs = {}

    def local_begin(q, k, slot=0):
        o, ticket = ix.search_batch_device_begin(q, k, stream=stream.cuda_stream, out=outs.get((slot, k, q.shape[0])))
        outs[(slot, k, q.shape[0])] = o
        return o, ticket

    sh = ShardedSearch(local_search, row_offset=rank * a.rows, local_begin=local_begin,
                       local_end=ix.search_batch_device_end, ticket_ok_ptr=ix.ticket_ok_ptr)

    def run_steps(n, depth=None):
        """n steps with up to --pipeline-depth searches in flight: the next batch is enqueued before the
        host waits for the previous one, so launch and wait latency are off the GPU's critical path.  Every
        step's result is complete and verified (search_end) inside the timed region."""
        depth = a.pipeline_depth if depth is None else depth
        last = None
        if depth <= 1:
            for _ in range(n):
                last = sh.search(d_q, a.k)
            return last
        pend = []
        for i in range(n):
            pend.append(sh.search_begin(d_q, a.k, slot=i % depth))
            if len(pend) >= depth:
                last = sh.search_end(pend.pop(0))
        while pend:
            last = sh.search_end(pend.pop(0))
        return last

    # host buffers of the end-to-end measurement: pinned (the contract's e2e) and pageable (what a caller that
    # flattens Vec<f32> queries hands over)
    h_q = torch.empty((a.batch, a.dim), dtype=torch.float32).pin_memory()
    h_q.copy_(d_q)
    h_q_np = h_q.numpy()
    h_q_pageable = np.array(h_q_np, copy=True)
    h_out = (torch.empty((a.batch, a.k), dtype=torch.int64).pin_memory(),
             torch.empty((a.batch, a.k), dtype=torch.float32).pin_memory(),
             torch.empty((a.batch, a.k), dtype=torch.float32).pin_memory(),
             torch.empty((a.batch,), dtype=torch.int32).pin_memory())

    def step_e2e(q_np=None):
        if world == 1:
            return ix.search_batch_arrays(h_q_np if q_np is None else q_np, a.k)  # host in, host out: the reference-facing call
        dq = h_q.to(dev, non_blocking=True)
        res = sh.search(dq, a.k)
        for hb, t in zip(h_out, res):          # pinned host buffers, one wait for all four copies
            hb.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h_out

    # ---- value: device-resident ---------------------------------------------------
    clocks = ClockSampler(world)
    if rank == 0:
        clocks.start()
    for _ in range(max(3, a.warmup)):
        sh.search(d_q, a.k)
    # untimed soak (~1 s) so that clocks settle and nvidia-smi gets samples under load
    t_soak = time.perf_counter()
    n_soak = torch.zeros(1, device=dev)
    while True:
        for _ in range(8):
            sh.search(d_q, a.k)
        n_soak.fill_(1.0 if time.perf_counter() - t_soak < 1.0 else 0.0)
        if world > 1:
            dist.broadcast(n_soak, src=0)
        if n_soak.item() == 0.0:
            break
    run_steps(3 * max(1, a.pipeline_depth))  # warm the pipelined form too (second workspace, exchange buffers, graphs)
    barrier()
    st0 = ix.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    timed_result = run_steps(a.steps)
    e1.record()
    barrier()
    ms_rank = e0.elapsed_time(e1)
    st1 = ix.stats()
    ms_total = max_over_ranks(ms_rank)
    ms_step = ms_total / a.steps
    value = a.batch * world * a.steps / (ms_total * 1e-3)
    timed_rows = timed_result[0].cpu().numpy()
    timed_score, timed_dist, timed_n = (t.cpu().numpy() for t in timed_result[1:])
    per_rank = None
    if world > 1:
        k_us = (st1["pass_kernel_ns"] - st0["pass_kernel_ns"]) * 1e-3 / max(1, st1["pass_kernel_launches"] - st0["pass_kernel_launches"])
        tl = torch.tensor([ms_rank / a.steps, k_us], device=dev, dtype=torch.float64)
        tl_all = torch.empty((world, 2), device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(tl_all.view(-1), tl)
        steps_ms = tl_all[:, 0].tolist()
        per_rank = {"step_ms_per_rank": [round(float(x), 4) for x in steps_ms],
                    "scan_us_per_rank": [round(float(x), 1) for x in tl_all[:, 1].tolist()],
                    "step_spread": (max(steps_ms) - min(steps_ms)) / max(steps_ms)}

    # ---- e2e: host buffers in and out ---------------------------------------------
    def time_e2e(fn, n):
        for _ in range(max(3, a.warmup)):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            res = fn()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        return a.batch * world * n / dt, res

    e2e_value, e2e_res = time_e2e(step_e2e, a.steps)
    e2e_single, e2e_pageable, e2e_callers = e2e_value, None, 1
    if world == 1:
        e2e_pageable, _ = time_e2e(lambda: step_e2e(h_q_pageable), a.steps)
        # the reference serves searches from many threads under read guards (serve.rs:101, routes.rs:906); the
        # library is re-entrant, so two callers overlap one call's copies and select step with the other's scan
        # -- the host-side counterpart of the two searches in flight of `value`
        e2e_callers = 2
        per = max(1, a.steps // e2e_callers)
        start = threading.Barrier(e2e_callers + 1)
        done = []

        def caller(i):
            torch.cuda.set_device(local_rank)
            q_i = h_q_np if i == 0 else np.array(h_q_np, copy=True)
            for _ in range(3):
                ix.search_batch_arrays(q_i, a.k)
            start.wait()
            for _ in range(per):
                ix.search_batch_arrays(q_i, a.k)
            done.append(time.perf_counter())
        ths = [threading.Thread(target=caller, args=(i,)) for i in range(e2e_callers)]
        for t in ths:
            t.start()
        start.wait()
        t0 = time.perf_counter()
        for t in ths:
            t.join()
        e2e_value = a.batch * per * e2e_callers / (max(done) - t0)
    clk = clocks.stop() if rank == 0 else None
    h2d = a.batch * a.dim * 4
    d2h = a.batch * a.k * (16 + 4 + 4) + a.batch * 4

    # ---- small-batch probe (B=1 interactive search, the HBM-bound streaming pass) ----
    # Default path: the pass streams the bf16 shadow of the rows (768 MB per query at 1M x 384) to nominate and
    # rescores exactly; `fp32_rows` is the same call with option stream_bf16=0 (1.536 GB per query).
    small, small_res = None, None
    if a.batch != 1 and not a.no_small_probe:
        q1 = d_q[:1].contiguous()
        n1 = 100

        def small_probe():
            o1 = None
            for _ in range(5):
                o1 = ix.search_batch_device(q1, a.k, stream=stream.cuda_stream, out=o1)
            torch.cuda.synchronize()
            s0 = ix.stats()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(n1):
                o1 = ix.search_batch_device(q1, a.k, stream=stream.cuda_stream, out=o1)
            ev1.record()
            torch.cuda.synchronize()
            s1 = ix.stats()
            res = tuple(t.cpu().numpy() for t in o1)
            for _ in range(5):  # warm-up like the device-resident loop above (the repeated call shape is recorded once)
                ix.search_batch_arrays(h_q_pageable[:1], a.k)
            t_host = time.perf_counter()
            for _ in range(n1):
                ix.search_batch_arrays(h_q_pageable[:1], a.k)
            host_qps = n1 / (time.perf_counter() - t_host)
            l1 = s1["pass_kernel_launches"] - s0["pass_kernel_launches"]
            ns1 = s1["pass_kernel_ns"] - s0["pass_kernel_ns"]
            if not (l1 and ns1):
                return None, res
            half = s1["queries_stream_bf16"] > s0["queries_stream_bf16"]
            # bytes of the rows one pass reads: the bf16 shadow (row padded to 64 elements) or the fp32 rows
            bytes1 = a.rows * ((a.dim + 63) // 64 * 64) * 2 if half else a.rows * ((a.dim + 3) // 4 * 4) * 4
            gbs = bytes1 / (ns1 * 1e-9 / l1) / 1e9
            return {"batch": 1, "rows_read_as": "bf16 shadow" if half else "fp32",
                    "queries_per_s": n1 / (ev0.elapsed_time(ev1) * 1e-3),
                    "ms_per_query": ev0.elapsed_time(ev1) / n1, "host_to_host_queries_per_s": host_qps,
                    "roofline": {"bound": "hbm", "kernel": "stream_scan_kernel", "achieved": gbs,
                                 "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                                 "frac_of_nominal_8TBs": gbs / 8000.0, "us_per_launch": ns1 * 1e-3 / l1,
                                 "traffic": ncu_traffic("stream_scan_kernel" if half else "stream_scan_kernel_fp32", 1),
                                 "algorithmic_bytes_per_launch": bytes1, "peak_source": pk["source"]}}, res

        small, small_res = small_probe()
        ix.set_option("stream_bf16", 0)
        try:
            fp32_rows, fp32_res = small_probe()
        finally:
            ix.set_option("stream_bf16", 1)
        if small is not None:
            small["fp32_rows"] = fp32_rows
            small["fp32_rows_same_result"] = bool(all(np.array_equal(x.view(np.uint8), y.view(np.uint8))
                                                      for x, y in zip(small_res, fp32_res)))

    # ---- auto-link cycle (BASELINE.json metric "auto-link pairs/s at 1/2/4/8 B200", configs[2] shape at a
    # bounded size): every new node searches the row-sharded corpus for its 100 nearest neighbours
    # (linker/auto_linker.rs:215-222), candidates with score >= 0.75 become links, at most 50 per node
    # (linker/rules.rs:50, auto_linker.rs:261).  N = 1: ONE call of cx_autolink_batch_device; N > 1: sharded
    # search + the same post-pass kernel on the merged lists.  Unit = scored (new node, corpus row) pair.
    def autolink_cycle_fn(index, rows_per_gpu, queries):
        out_al = {}

        def local_al(q, k):
            out_al[0] = index.search_batch_device(q, k, stream=stream.cuda_stream, out=out_al.get(0))
            return out_al[0]

        def local_al_begin(q, k, slot=0):
            out_al[0], ticket = index.search_batch_device_begin(q, k, stream=stream.cuda_stream, out=out_al.get(0))
            return out_al[0], ticket

        sh_al = ShardedSearch(local_al, row_offset=rank * rows_per_gpu, local_begin=local_al_begin,
                              local_end=index.search_batch_device_end, ticket_ok_ptr=index.ticket_ok_ptr)
        bufs = {}

        def cycle():
            if world == 1:
                res, bufs[0] = index.autolink_batch_device(queries, 100, 0.75, 50, stream=stream.cuda_stream, bufs=bufs.get(0))
                return res
            return sh_al.autolink(queries, None, 100, 0.75, 50)
        return cycle

    def time_cycles(cycle, n_cyc, index):
        for _ in range(2):
            res = cycle()
        barrier()
        sa0 = index.stats()
        ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea0.record()
        for _ in range(n_cyc):
            res = cycle()
            links = res[2].sum()
        ea1.record()
        barrier()
        sa1 = index.stats()
        return max_over_ranks(ea0.elapsed_time(ea1)) / n_cyc, int(links.item()), sa0, sa1

    autolink = None
    if a.autolink_new > 0:
        nq = a.autolink_new
        if world > 1:
            dist.broadcast(q_al, src=0)
        q_al = q_al.contiguous()
        ms_cyc, links, sa0, sa1 = time_cycles(autolink_cycle_fn(ix, a.rows, q_al), 3, ix)
        tf = 2.0 * a.dim * float(nq) * a.rows / (ms_cyc * 1e-3) / 1e12  # per GPU
        autolink = {"metric": "scored pairs/s", "value": float(nq) * a.rows * world / (ms_cyc * 1e-3), "unit": "pairs/s",
                    "new_nodes_per_cycle": nq, "rows_per_gpu": a.rows, "k": 100, "threshold": 0.75, "max_edges": 50,
                    "ms_per_cycle": ms_cyc, "link_candidates": links,
                    "call": "cx_autolink_batch_device" if world == 1 else "sharded search + cx_autolink_filter_device",
                    "tflops_per_gpu_whole_cycle": tf, "frac_of_sustained_bf16_whole_cycle": tf / pk["bf16_tflops_sustained"],
                    "fallbacks": sa1["fallbacks"] - sa0["fallbacks"]}

    # ---- N > 1: rank 0 rebuilds the whole corpus on the host for the parity check of the merged result
    parity = {"bar": "ids equal, score and distance bits equal (oracle/cortex_oracle.c on the same corpus and queries)"}
    cpu = None
    if rank == 0:
        from oracle.binding import OracleIndex

        n_par = 16 if world == 1 else 8
        o_ix = OracleIndex(a.dim, faithful_copy=True)
        o_ix.insert_batch(ids_for(a.rows), corpus_np)
        for r in range(1, world):  # the other ranks' shards, regenerated from their seeds
            c_r = shard_corpus(r).cpu().numpy()
            o_ix.insert_batch(ids_for(a.rows, r * a.rows), c_r)
            del c_r
        q_np = q_all.cpu().numpy()
        from oracle.binding import max_threads

        o_ix._L.cxo_set_faithful_copy(o_ix._h, 0)  # the checker does not need the per-pair clone
        o_res = o_ix.search_batch(q_np[:n_par], a.k, n_threads=max_threads())
        parity["queries"] = n_par
        parity["corpus_rows"] = a.rows * world
        parity["identical"] = compare_with_oracle(o_res, timed_rows, timed_score, timed_dist, timed_n)
        parity["what"] = "the last step of the timed device-resident loop" + (" (merged over all ranks)" if world > 1 else "")
        if world == 1:
            ids_e, sc_e, di_e, n_e = e2e_res
            parity["e2e_identical"] = compare_with_oracle(o_res, rows_from_ids(ids_e), sc_e, di_e, n_e)
        else:
            parity["e2e_identical"] = compare_with_oracle(o_res, *[t.numpy() for t in e2e_res])
        if small_res is not None and world == 1:
            o1_res = tuple(x[:1] for x in o_res)
            parity["small_batch_identical"] = compare_with_oracle(o1_res, *small_res)
        o_ix._L.cxo_set_faithful_copy(o_ix._h, 1)

        # ---- CPU baselines (rank 0, N=1 only) ---------------------------------------
        if world == 1 and not a.no_cpu_baseline:
            qps, cores, ms, sample = cpu_reference_leg(o_ix, q_np, a.k, 2, 1, a.cpu_queries)
            cpu = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                   "optimised_courtesy": cpu_optimised_leg(corpus_np, q_np, a.k)}
            if not a.no_hnsw:
                ex10 = rows_from_ids(ix.search_batch_arrays(q_np[:64], 10)[0])
                ex100 = rows_from_ids(ix.search_batch_arrays(q_np[:64], 100)[0])
                cpu["hnsw"] = cpu_hnsw_leg(corpus_np, q_np, ex10, ex100, a.hnsw_rows)
        del o_ix
    host_barrier()

    # ---- strong scaling: BASELINE.json configs[2] as named -------------------------------------------------
    strong = None
    if a.strong_rows > 0:
        rows_s = a.strong_rows // world
        ixs = GpuVectorIndex(a.dim, device=local_rank)
        ixs.reserve(rows_s)
        chunk = 250_000  # the corpus is the same 250k-row chunks whatever N is: rank r holds chunks r*rows_s/chunk ...
        q_src = None
        for s0 in range(0, rows_s, chunk):
            n = min(chunk, rows_s - s0)
            g_chunk = (rank * rows_s + s0) // chunk
            # clusters of ~1024 rows: every shard of an 8-GPU run still holds more than k = 100 members of each
            c = make_corpus_torch(n, a.dim, SEED + 31 * g_chunk + 104729, dev, cluster_seed=SEED + 3,
                                  n_clusters=a.strong_rows // 1024)
            ixs.insert_batch_device(ids_for(n, rank * rows_s + s0), c)
            if s0 == 0 and rank == 0:
                q_src = c[:min(n, 65536)].clone()
            del c
        torch.cuda.empty_cache()
        nq_s = a.strong_new
        q_new = torch.empty((nq_s, a.dim), device=dev, dtype=torch.float32)
        if rank == 0:
            g = torch.Generator(device=dev)
            g.manual_seed(7)
            idx = torch.randint(0, q_src.shape[0], (nq_s,), generator=g, device=dev)
            q_new = q_src[idx] + torch.randn((nq_s, a.dim), generator=g, device=dev) * 0.03
            q_new /= q_new.norm(dim=1, keepdim=True)
            q_new = q_new.contiguous()
        if world > 1:
            dist.broadcast(q_new, src=0)
        # the cycle in batches of one tensor-pass launch group (148 x 128 new nodes)
        QB = 148 * 128
        chunks = [q_new[s0:s0 + QB].contiguous() for s0 in range(0, nq_s, QB)]
        ixs.set_option("profile", 1)
        if world == 1:
            cyc = [autolink_cycle_fn(ixs, rows_s, qc) for qc in chunks]

            def cycle_all():
                res = None
                for c in cyc:
                    res = c()
                return res
        else:
            # two batches in flight: the exchange + merge + post-pass of one overlaps the scan of the next
            outs_s = {}

            def ls(q, k):
                outs_s[(0, q.shape[0])] = ixs.search_batch_device(q, k, stream=stream.cuda_stream, out=outs_s.get((0, q.shape[0])))
                return outs_s[(0, q.shape[0])]

            def lsb(q, k, slot=0):
                key = (slot, q.shape[0])
                outs_s[key], t = ixs.search_batch_device_begin(q, k, stream=stream.cuda_stream, out=outs_s.get(key))
                return outs_s[key], t

            sh_s = ShardedSearch(ls, row_offset=rank * rows_s, local_begin=lsb, local_end=ixs.search_batch_device_end,
                                 ticket_ok_ptr=ixs.ticket_ok_ptr)

            def cycle_all():
                pend, res = None, None
                for i, qc in enumerate(chunks):
                    p = sh_s.autolink_begin(qc, 100, slot=i % 2)
                    if pend is not None:
                        res = sh_s.autolink_end(pend, None, 0.75, 50)
                    pend = p
                return sh_s.autolink_end(pend, None, 0.75, 50)
        res = cycle_all()
        barrier()
        ss0 = ixs.stats()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_cyc = 2
        es0.record()
        for _ in range(n_cyc):
            res = cycle_all()
        es1.record()
        barrier()
        ss1 = ixs.stats()
        ms_c = max_over_ranks(es0.elapsed_time(es1)) / n_cyc
        pairs = float(nq_s) * float(rows_s * world)
        tf = 2.0 * a.dim * float(nq_s) * rows_s / (ms_c * 1e-3) / 1e12
        scan_ms = (ss1["pass_kernel_ns"] - ss0["pass_kernel_ns"]) * 1e-6 / n_cyc
        strong = {"workload": f"auto-link cycle: {nq_s} new nodes x {rows_s * world}-row {a.dim}-d corpus, k=100, threshold 0.75, "
                              f"max 50 links per node, rows split over {world} GPU(s) (BASELINE.json configs[2])",
                  "scaling": "strong", "metric": "scored pairs/s", "value": pairs / (ms_c * 1e-3), "unit": "pairs/s",
                  "ms_per_cycle": ms_c, "rows_per_gpu": rows_s, "new_nodes": nq_s,
                  "tflops_per_gpu_whole_cycle": tf, "frac_of_sustained_bf16_whole_cycle": tf / pk["bf16_tflops_sustained"],
                  "scan_ms_per_cycle": scan_ms, "scan_share_of_cycle": scan_ms / ms_c,
                  "fallbacks": ss1["fallbacks"] - ss0["fallbacks"]}
        del ixs
        torch.cuda.empty_cache()

    # ---- configs[3] as named: 50M x 1024-d bf16-valued corpus over 8 GPUs, B = 256, top-100 ----------------
    cfg4 = None
    if a.cfg4 == "on" or (a.cfg4 == "auto" and world == 8):
        d4, rows4, B4, k4 = 1024, 50_000_000 // 8, 256, 100
        ix4 = GpuVectorIndex(d4, device=local_rank)
        ix4.reserve(rows4)
        q4 = None
        for s0 in range(0, rows4, 625_000):
            n = min(625_000, rows4 - s0)
            c = make_corpus_torch(n, d4, SEED + 17 * (s0 // 625_000) + 7919 * rank, dev, cluster_seed=SEED + 4,
                                  n_clusters=50_000_000 // 1024)
            c = c.to(torch.bfloat16).to(torch.float32)  # the corpus is bf16: rows hold bf16 values
            ix4.insert_batch_device(ids_for(n, rank * rows4 + s0), c)
            if s0 == 0:
                q4 = make_queries_torch(c, B4, SEED)
            del c
        torch.cuda.empty_cache()
        if world > 1:
            dist.broadcast(q4, src=0)
        ix4.set_option("profile", 1)
        o4 = {}

        def l4(q, k):
            o4[0] = ix4.search_batch_device(q, k, stream=stream.cuda_stream, out=o4.get(0))
            return o4[0]

        def l4b(q, k, slot=0):
            o4[slot], t = ix4.search_batch_device_begin(q, k, stream=stream.cuda_stream, out=o4.get(slot))
            return o4[slot], t

        sh4 = ShardedSearch(l4, row_offset=rank * rows4, local_begin=l4b, local_end=ix4.search_batch_device_end,
                            ticket_ok_ptr=ix4.ticket_ok_ptr)
        for _ in range(4):
            sh4.search(q4, k4)
        barrier()
        s40 = ix4.stats()
        e40, e41 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n4 = 20
        e40.record()
        pend = []
        t_b = t_e = 0.0
        for i in range(n4):
            tb = time.perf_counter()
            pend.append(sh4.search_begin(q4, k4, slot=i % 2))
            t_b += time.perf_counter() - tb
            if len(pend) >= 2:
                te = time.perf_counter()
                sh4.search_end(pend.pop(0))
                t_e += time.perf_counter() - te
        while pend:
            sh4.search_end(pend.pop(0))
        e41.record()
        print(f"[cfg4 rank {rank}] host ms per step: begin {1e3 * t_b / n4:.3f} end {1e3 * t_e / n4:.3f}", file=sys.stderr)
        barrier()
        s41 = ix4.stats()
        ms4 = max_over_ranks(e40.elapsed_time(e41)) / n4
        us4 = (s41["pass_kernel_ns"] - s40["pass_kernel_ns"]) * 1e-3 / max(1, s41["pass_kernel_launches"] - s40["pass_kernel_launches"])
        tf4 = 2.0 * d4 * B4 * rows4 / (us4 * 1e-6) / 1e12
        gb4 = rows4 * d4 * 2 / (us4 * 1e-6) / 1e9
        cfg4 = {"workload": f"search: {rows4 * world} x {d4}-d bf16-valued corpus over {world} GPU(s), batch {B4}, top-{k4} "
                            f"(BASELINE.json configs[3])",
                "queries_per_s": B4 / (ms4 * 1e-3), "ms_per_step": ms4, "rows_per_gpu": rows4,
                "scan_us_per_gpu": us4, "tflops_per_gpu": tf4, "frac_of_sustained_bf16": tf4 / pk["bf16_tflops_sustained"],
                "shadow_stream_gbs_per_gpu": gb4, "frac_of_hbm": gb4 / pk["hbm_gbs"],
                "store_gb_per_gpu": rows4 * (d4 * 4 + d4 * 2 + 32) / 1e9, "fallbacks": s41["fallbacks"] - s40["fallbacks"]}
        del ix4, sh4
        torch.cuda.empty_cache()

    # ---- N > 1: ONE process driving all N GPUs behind the C ABI (cx_index_create_sharded) ---------------------
    single = None
    if world > 1 and not a.no_single_process:
        barrier()
        host_barrier()
        if rank == 0:
            try:
                ixm = GpuVectorIndex(a.dim, devices=list(range(world)))
                for r in range(world):
                    c_r = shard_corpus(r)
                    ixm.insert_batch_device(ids_for(a.rows, r * a.rows), c_r)
                    del c_r
                torch.cuda.empty_cache()
                om = [None, None]
                idm = [torch.zeros((a.batch, a.k, 16), dtype=torch.uint8, device=dev) for _ in range(2)]

                def m_steps(n):
                    pend = []
                    for i in range(n):
                        sl = i % 2
                        om[sl], t = ixm.search_batch_device_begin(d_q, a.k, stream=stream.cuda_stream, out=om[sl], ids_out=idm[sl])
                        pend.append(t)
                        if len(pend) >= 2:
                            ixm.search_batch_device_end(pend.pop(0))
                    while pend:
                        ixm.search_batch_device_end(pend.pop(0))
                m_steps(12)
                torch.cuda.synchronize()
                m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                m0.record()
                m_steps(a.steps)
                m1.record()
                torch.cuda.synchronize()
                ms_m = m0.elapsed_time(m1) / a.steps
                last = (a.steps - 1) % 2
                rows_m = rows_from_ids(idm[last].cpu().numpy())
                ok_m = compare_with_oracle(o_res, rows_m, om[last][1].cpu().numpy(), om[last][2].cpu().numpy(),
                                           om[last][3].cpu().numpy())
                t0 = time.perf_counter()
                for _ in range(a.steps):
                    ixm.search_batch_arrays(h_q_pageable, a.k)
                e2e_m = a.batch * a.steps / (time.perf_counter() - t0)
                stm = ixm.stats()
                single = {"what": f"one process, {world} GPUs behind cx_index_create_sharded; local lists are stored into device 0 "
                                  f"by the producing GPUs and merged there (no NCCL)",
                          "ms_per_step": ms_m, "shard_scans_per_s": a.batch * world / (ms_m * 1e-3),
                          "queries_per_s": a.batch / (ms_m * 1e-3), "host_to_host_queries_per_s": e2e_m,
                          "parity_identical": ok_m, "graph_launches": stm["graph_launches"]}
                del ixm
            except Exception as e:  # noqa: BLE001
                single = {"failed": repr(e)[:300]}
        host_barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the scan pass) --------------------------
    launches = st1["pass_kernel_launches"] - st0["pass_kernel_launches"]
    ns = st1["pass_kernel_ns"] - st0["pass_kernel_ns"]
    ld = (a.dim + 3) // 4 * 4
    alg_bytes = a.rows * ld * 4
    roof = None
    used_tensor = (st1["queries_tensor"] - st0["queries_tensor"]) > 0
    if launches and ns and not used_tensor:
        sec = ns * 1e-9 / launches
        achieved = alg_bytes / sec / 1e9
        roof = {"bound": "hbm", "kernel": "stream_scan_kernel", "achieved": achieved, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": ncu_traffic("stream_scan_kernel", a.batch),
                "peak_source": pk["source"], "us_per_launch": sec * 1e6, "launches": launches,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_share_of_step": (ns * 1e-6) / ms_rank}
    elif launches and ns:
        # tensor pass: 2*D flops per scored (query,row) pair; the kernel runs inside a
        # seconds-long loop, so the sustained cuBLAS figure is the denominator (burst also given)
        sec = ns * 1e-9 / launches
        q_per_launch = a.batch * a.steps / launches
        flops = 2.0 * a.dim * q_per_launch * a.rows
        achieved = flops / sec / 1e12
        hbm_bytes = a.rows * ((a.dim + 63) // 64 * 64) * 2  # bf16 shadow streamed once per launch
        roof = {"bound": "tensor", "kernel": "tensor_scan_kernel", "achieved": achieved,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops_sustained"], "frac_of_burst": achieved / pk["bf16_tflops"],
                "traffic": ncu_traffic("tensor_scan_kernel", a.batch), "peak_source": pk["source"],
                "us_per_launch": sec * 1e6, "launches": launches,
                "algorithmic_flops_per_launch": flops,
                "launch_note": "one launch = one scan of the shard: the cut-off bootstrap (sampled tiles + select), the "
                               "phase launches of tensor_scan_kernel and the tau_refine kernels between them; the CUDA "
                               "events bracket all of it, inside the timed region",
                "hbm_floor_us": hbm_bytes / (pk["hbm_gbs"] * 1e9) * 1e6,
                "hbm_gbs_of_shadow_stream": hbm_bytes / sec / 1e9,
                "kernel_share_of_step": (ns * 1e-6) / ms_rank}

    cfg = workload_config(a, world)
    cfg["pipeline_depth"] = a.pipeline_depth
    line = {
        "metric": "queries/s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "queries_per_s": a.batch / (ms_step * 1e-3),
        "pairs_per_s": float(a.batch) * a.rows * world / (ms_step * 1e-3),
        "corpus_rows_total": a.rows * world,
        "parity": parity,
        "roofline": roof, "small_batch": small, "autolink": autolink, "strong_scaling": strong, "cfg4": cfg4,
        "single_process": single, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "buffers": "pinned host memory" if e2e_callers == 1 else "caller 0 pinned, caller 1 pageable",
                "concurrent_callers": e2e_callers, "single_caller_value": e2e_single, "pageable_value": e2e_pageable},
        "gpu_launches": st1["kernel_launches"] - st0["kernel_launches"],
        "graph_launches": st1["graph_launches"] - st0["graph_launches"],
        "per_rank": per_rank, "host_cpus": os.cpu_count(),
        "clocks": clk,
        "paths": {"stream": st1["queries_stream"] - st0["queries_stream"],
                  "tensor": st1["queries_tensor"] - st0["queries_tensor"],
                  "exact": st1["queries_exact"] - st0["queries_exact"],
                  "fallbacks": st1["fallbacks"] - st0["fallbacks"]},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
