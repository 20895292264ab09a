#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configs beyond bench.py's headline line.

Writes one JSON document (default profiles/results_r2.json) with a section per config:
  cfg1  10k x 384, 100 queries, top-10: GPU vs CPU oracle, ids/score bits compared
  cfg2  1M x 384, query batch sweep 1..1024, top-10: queries/s, kernel roofline per batch
  cfg3  auto-link cycle: B new nodes x N-row corpus, k=100 (+ threshold 0.75): scored pairs/s
  dedup the dedup scanner's self-join over 1M rows at threshold 0.92
  hnsw  recall@10 / @100 of the reference's approximate path (restated on the CPU) against the exact result
  cfg5  streaming ingest: 256-node batches searched (k=100) then appended to a growing corpus,
        p50/p99 latency per batch
  cfg4  one GPU's shard (6.25M rows) of the 50M x 1024-d corpus, batch 256, top-100
All timing is CUDA events on the caller's stream around the public device API.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (synthetic data + peaks)


def ids_for(n, start=0):
    ids = np.zeros((n, 16), np.uint8)
    ids[:, 8:] = (np.arange(n, dtype=np.uint64) + start).astype(">u8").view(np.uint8).reshape(-1, 8)
    return ids


def timed(fn, steps, warmup, torch):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps  # ms per call


def cfg1(torch, out):
    from cortex_b200 import GpuVectorIndex, synth
    from oracle.binding import OracleIndex, max_threads

    corpus = synth.make_corpus(10_000, 384, zero_row=True)
    Q = synth.make_queries(corpus, 100)
    ids = synth.make_ids(10_000)
    g = GpuVectorIndex(384)
    g.insert_batch(ids, corpus)
    o = OracleIndex(384, faithful_copy=True)
    o.insert_batch(ids, corpus)
    t0 = time.perf_counter()
    oi, osc, od, _, on = o.search_batch(Q, 10, n_threads=max_threads())
    t_cpu = time.perf_counter() - t0
    gi, gs, gd, gn = g.search_batch_arrays(Q, 10)
    t0 = time.perf_counter()
    for _ in range(20):
        g.search_batch_arrays(Q, 10)
    t_gpu = (time.perf_counter() - t0) / 20
    same = (np.array_equal(gi, oi) and np.array_equal(gn, on)
            and bool(np.all((gs.view(np.uint32) == osc.view(np.uint32)) | (np.isnan(gs) & np.isnan(osc)))))
    out["cfg1"] = {"workload": "10k x 384, 100 queries, top-10", "ids_and_score_bits_identical_to_oracle": same,
                   "gpu_ms_per_batch_host_to_host": t_gpu * 1e3, "gpu_queries_per_s": 100 / t_gpu,
                   "cpu_oracle_ms_per_batch": t_cpu * 1e3, "cpu_queries_per_s": 100 / t_cpu,
                   "cpu_threads": max_threads(), "paths": g.stats()}


def cfg2(torch, out, rows):
    from cortex_b200 import GpuVectorIndex

    dev = torch.device("cuda", 0)
    corpus = bench.make_corpus_torch(rows, 384, bench.SEED, dev)
    q_all = bench.make_queries_torch(corpus, 1024, bench.SEED)
    ix = GpuVectorIndex(384)
    ix.reserve(rows)
    ix.insert_batch_device(ids_for(rows), corpus)
    del corpus
    torch.cuda.empty_cache()
    ix.set_option("profile", 1)
    pk = bench.peaks()
    s = torch.cuda.current_stream().cuda_stream
    res = []
    for B in (1, 2, 3, 4, 5, 8, 16, 32, 64, 128, 256, 512, 1024):
        dq = q_all[:B].contiguous()
        hq = dq.cpu().numpy()
        buf = [None]

        def dev_step():
            buf[0] = ix.search_batch_device(dq, 10, stream=s, out=buf[0])

        steps = 50 if B <= 64 else 20
        dev_step()
        st0 = ix.stats()
        ms = timed(dev_step, steps, 3, torch)
        st1 = ix.stats()
        for _ in range(3):
            ix.search_batch_arrays(hq, 10)
        t0 = time.perf_counter()
        for _ in range(steps):
            ix.search_batch_arrays(hq, 10)
        ms_host = (time.perf_counter() - t0) / steps * 1e3
        launches = st1["pass_kernel_launches"] - st0["pass_kernel_launches"]
        ns = st1["pass_kernel_ns"] - st0["pass_kernel_ns"]
        tensor = st1["queries_tensor"] > st0["queries_tensor"]
        half = st1["queries_stream_bf16"] > st0["queries_stream_bf16"]
        us = ns * 1e-3 / max(1, launches)
        row = {"batch": B, "queries_per_s": B / (ms * 1e-3), "ms_per_batch": ms,
               "host_to_host_queries_per_s": B / (ms_host * 1e-3),
               "pass": "tensor" if tensor else ("stream (bf16 shadow)" if half else "stream (fp32 rows)"),
               "scan_kernel_us": us, "fallbacks": st1["fallbacks"] - st0["fallbacks"]}
        if tensor:
            tf = 2.0 * 384 * B * rows / (us * 1e-6) / 1e12
            row["tflops"] = tf
            row["frac_of_sustained_bf16"] = tf / pk["bf16_tflops_sustained"]
            row["shadow_stream_gbs"] = rows * 384 * 2 / (us * 1e-6) / 1e9
            row["frac_of_hbm_for_shadow_stream"] = row["shadow_stream_gbs"] / pk["hbm_gbs"]
        else:
            gbs = rows * 384 * (2 if half else 4) / (us * 1e-6) / 1e9  # the rows (shadow or fp32) are read once per launch
            row["hbm_gbs"] = gbs
            row["frac_of_measured_hbm"] = gbs / pk["hbm_gbs"]
        res.append(row)
    out["cfg2"] = {"workload": f"{rows} x 384 fp32, top-10, query batch sweep", "rows": res}
    return ix, q_all


def cfg3(torch, out, rows, n_new):
    """auto-link cycle: every new node searches the corpus for its 100 nearest neighbours
    (linker/auto_linker.rs:215-222), candidates with score >= 0.75 become links (rules.rs:50)."""
    from cortex_b200 import GpuVectorIndex

    dev = torch.device("cuda", 0)
    ix = GpuVectorIndex(384)
    ix.reserve(rows)
    chunk = 2_000_000
    for s0 in range(0, rows, chunk):
        n = min(chunk, rows - s0)
        c = bench.make_corpus_torch(n, 384, bench.SEED + 31 * (s0 // chunk), dev)
        ix.insert_batch_device(ids_for(n, s0), c)
        if s0 == 0:
            q_src = c[:max(n_new, 1024)].clone() if n >= n_new else c.clone()
        del c
    torch.cuda.empty_cache()
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    reps = (n_new + q_src.shape[0] - 1) // q_src.shape[0]
    Q = q_src.repeat(reps, 1)[:n_new].clone()
    Q += torch.randn(Q.shape, generator=g, device=dev) * 0.03
    Q /= Q.norm(dim=1, keepdim=True)
    Q = Q.contiguous()
    s = torch.cuda.current_stream().cuda_stream
    buf = [None]

    def step():
        buf[0] = ix.search_batch_device(Q, 100, stream=s, out=buf[0])

    step()
    st0 = ix.stats()
    ms = timed(step, 2, 0, torch)
    st1 = ix.stats()
    rows_t, score, dist, n = buf[0]
    links = int((score >= 0.75).sum().item())
    pairs = float(n_new) * rows
    tf = 2.0 * 384 * pairs / (ms * 1e-3) / 1e12
    pk = bench.peaks()
    out["cfg3"] = {"workload": f"auto-link cycle: {n_new} new nodes x {rows}-row corpus, k=100, threshold 0.75",
                   "ms_per_cycle": ms, "scored_pairs_per_s": pairs / (ms * 1e-3), "tflops": tf,
                   "frac_of_sustained_bf16": tf / pk["bf16_tflops_sustained"],
                   "link_candidates_at_0.75": links,
                   "tensor_queries": st1["queries_tensor"] - st0["queries_tensor"],
                   "stream_retries": st1["queries_stream"] - st0["queries_stream"],
                   "exact_fallbacks": st1["queries_exact"] - st0["queries_exact"]}


def cfg5(torch, out, final_rows):
    """streaming ingest (BASELINE.json configs[4]): 256-node batches; each batch runs the auto-link scan step
    (k=100, threshold 0.75, at most 50 links per node: ONE call of cx_autolink_batch_device) against the corpus so
    far and is then appended.  Nothing is reserved: the store grows in place on the way to `final_rows`.  The
    latency of EVERY batch (scan + append, host clock around a device synchronisation) enters the percentiles;
    the data is generated on the device, a million rows at a time, outside the timed region."""
    from cortex_b200 import GpuVectorIndex

    dev = torch.device("cuda", 0)
    ix = GpuVectorIndex(384)
    s = torch.cuda.current_stream().cuda_stream
    seed_rows = 4096
    c = bench.make_corpus_torch(seed_rows, 384, bench.SEED + 5, dev)
    ix.insert_batch_device(ids_for(seed_rows), c)
    pool_rows = 1 << 20  # fresh rows are generated a million at a time (no row is ever inserted twice)
    pool, pool_at = None, pool_rows
    n_rows = seed_rows
    bufs = None
    lat, at_rows, links = [], [], 0
    b = 0
    grow0 = ix.stats()["grow_events"]
    while n_rows + 256 <= final_rows:
        if pool_at + 256 > pool_rows:
            pool = bench.make_corpus_torch(pool_rows, 384, bench.SEED + 6 + b, dev)
            pool = pool[torch.randperm(pool_rows, device=dev)].contiguous()  # arrival order is not cluster order
            pool_at = 0
        rows = pool[pool_at:pool_at + 256]
        pool_at += 256
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, bufs = ix.autolink_batch_device(rows, 100, 0.75, 50, stream=s, bufs=bufs)
        ix.insert_batch_device(ids_for(256, n_rows), rows)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
        at_rows.append(n_rows)
        n_rows += 256
        b += 1
        if b % 2000 == 0:
            links += int(res[2].sum().item())
    lat = np.asarray(lat)
    at_rows = np.asarray(at_rows)
    st = ix.stats()
    bands = {}
    for lo, hi in ((0, 100_000), (100_000, 500_000), (500_000, 1_000_000), (1_000_000, 2_000_000),
                   (2_000_000, 3_500_000), (3_500_000, 5_000_001)):
        m = (at_rows >= lo) & (at_rows < hi)
        if m.any():
            v = lat[m]
            bands[f"{lo}-{hi}"] = {"p50": float(np.percentile(v, 50)), "p99": float(np.percentile(v, 99)),
                                   "max": float(v.max()), "n": int(m.sum())}
    out["cfg5"] = {"workload": f"256-node batches: auto-link scan (k=100, threshold 0.75, cap 50) then append; corpus grows "
                               f"from {seed_rows} to {n_rows} rows with nothing reserved",
                   "batches": b, "total_s": float(lat.sum() * 1e-3), "batches_per_s": b / float(lat.sum() * 1e-3),
                   "latency_ms_all_batches": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                                              "p999": float(np.percentile(lat, 99.9)), "max": float(lat.max())},
                   "latency_ms_by_corpus_rows": bands,
                   "grow_events": st["grow_events"] - grow0, "in_place_growth": st["in_place_growth"],
                   "capacity_rows": st["capacity_rows"], "link_candidates_sampled": links, "paths": st}


def cfg4(torch, out, rows, batch):
    """large-embedding search: one GPU's shard of the 50M x 1024-d corpus (50M / 8 = 6.25M rows), top-100,
    batch 256.  The tensor pass scans the bf16 copy (rows x 1024 x 2 bytes per batch); the fp32 rows stay
    resident for the exact rescoring.  At B=256 this sits at the HBM / tensor ridge."""
    from cortex_b200 import GpuVectorIndex

    dev = torch.device("cuda", 0)
    ix = GpuVectorIndex(1024)
    ix.reserve(rows)
    chunk = 500_000
    q = None
    for s0 in range(0, rows, chunk):
        n = min(chunk, rows - s0)
        c = bench.make_corpus_torch(n, 1024, bench.SEED + 41 * (s0 // chunk), dev)
        ix.insert_batch_device(ids_for(n, s0), c)
        if q is None:
            q = bench.make_queries_torch(c, batch, bench.SEED + 4)
        del c
    torch.cuda.empty_cache()
    ix.set_option("profile", 1)
    s = torch.cuda.current_stream().cuda_stream
    buf = [None]

    def step():
        buf[0] = ix.search_batch_device(q, 100, stream=s, out=buf[0])

    step()
    st0 = ix.stats()
    ms = timed(step, 10, 2, torch)
    st1 = ix.stats()
    pk = bench.peaks()
    us = (st1["pass_kernel_ns"] - st0["pass_kernel_ns"]) * 1e-3 / max(1, st1["pass_kernel_launches"] - st0["pass_kernel_launches"])
    shadow_bytes = rows * 1024 * 2
    tf = 2.0 * 1024 * batch * rows / (us * 1e-6) / 1e12
    out["cfg4"] = {"workload": f"one shard of cfg4: {rows} x 1024-d (bf16 copy scanned, fp32 kept for rescoring), "
                               f"batch {batch}, top-100", "ms_per_batch": ms, "queries_per_s": batch / (ms * 1e-3),
                   "scan_us": us, "shadow_stream_gbs": shadow_bytes / (us * 1e-6) / 1e9,
                   "frac_of_measured_hbm": shadow_bytes / (us * 1e-6) / 1e9 / pk["hbm_gbs"], "tflops": tf,
                   "frac_of_sustained_bf16": tf / pk["bf16_tflops_sustained"],
                   "fallbacks": st1["fallbacks"] - st0["fallbacks"],
                   "tensor_queries": st1["queries_tensor"] - st0["queries_tensor"]}


def dedup(torch, out, rows):
    """dedup self-join (linker/dedup.rs:65-127): every node's search_threshold(0.92) as one scan of the
    upper triangle of E.E^T; unit of work = unordered (node, node) pair scored."""
    from cortex_b200 import GpuVectorIndex

    dev = torch.device("cuda", 0)
    ix = GpuVectorIndex(384)
    ix.reserve(rows)
    c = bench.make_corpus_torch(rows, 384, bench.SEED + 77, dev)
    ix.insert_batch_device(ids_for(rows), c)
    del c
    torch.cuda.empty_cache()
    ix.dedup_scan(0.92, per_node_cap=64, max_pairs=4 << 20)  # warm-up (workspace allocation)
    t0 = time.perf_counter()
    a, b, sc, total = ix.dedup_scan(0.92, per_node_cap=64, max_pairs=4 << 20)
    dt = time.perf_counter() - t0
    pairs = rows * (rows - 1) / 2.0
    pk = bench.peaks()
    tf = 2.0 * 384 * pairs / dt / 1e12
    out["dedup"] = {"workload": f"dedup self-join over {rows} x 384 rows, threshold 0.92, host call (ids out)",
                    "seconds": dt, "unordered_pairs_scored_per_s": pairs / dt, "tflops_upper_triangle": tf,
                    "frac_of_sustained_bf16": tf / pk["bf16_tflops_sustained"], "duplicate_pairs": int(total),
                    "pairs_returned": int(len(a)), "paths": ix.stats()}


def hnsw(torch, out, rows, n_queries):
    """The reference's approximate path (HnswIndex::search after rebuild, index.rs:342-373) restated on the
    CPU (oracle/hnsw_oracle.c -- parity unpinned, parameters unverified): recall@k against the exact
    result, which is what this library returns."""
    from cortex_b200 import GpuVectorIndex
    from oracle.binding import OracleHnsw, max_threads

    corpus = bench.make_corpus_torch(rows, 384, bench.SEED + 9, "cpu").numpy()
    Q = bench.make_queries_torch(torch.from_numpy(corpus), n_queries, bench.SEED + 9).numpy()
    g = GpuVectorIndex(384)
    g.insert_batch(ids_for(rows), corpus)
    t0 = time.perf_counter()
    hn = OracleHnsw(corpus, n_threads=max_threads())
    t_build = time.perf_counter() - t0
    res = {"workload": f"{rows} x 384, {n_queries} queries; HNSW restated with M=32, ef_construction=100, "
                       f"ef_search=100 (instant-distance defaults from memory, unverified)",
           "build_s": t_build, "build_threads": max_threads()}
    for k in (10, 100):
        ids, sc, di, n = g.search_batch_arrays(Q, k)
        exact_rows = ids[:, :, 8:].copy().view(">u8").reshape(n_queries, k).astype(np.int64)
        hit = 0
        t0 = time.perf_counter()
        got = [hn.search(q, k)[0] for q in Q]
        dt = time.perf_counter() - t0
        for b in range(n_queries):
            hit += len(set(int(x) for x in got[b]) & set(int(x) for x in exact_rows[b, :int(n[b])]))
        res[f"recall@{k}"] = hit / float(n.sum())
        res[f"cpu_hnsw_queries_per_s_k{k}"] = n_queries / dt
        res[f"results_returned_per_query_k{k}"] = float(np.mean([len(x) for x in got]))
    out["hnsw"] = res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "results_r2.json"))
    ap.add_argument("--only", default="1,2,3,4,dedup,hnsw,5")
    ap.add_argument("--cfg2-rows", type=int, default=1_000_000)
    ap.add_argument("--cfg3-rows", type=int, default=10_000_000)
    ap.add_argument("--cfg3-new", type=int, default=100_000)
    ap.add_argument("--cfg5-rows", type=int, default=5_000_000)
    ap.add_argument("--cfg4-rows", type=int, default=6_250_000)
    ap.add_argument("--dedup-rows", type=int, default=1_000_000)
    ap.add_argument("--hnsw-rows", type=int, default=10_000)
    a = ap.parse_args()
    import torch

    torch.cuda.set_device(0)
    out = {}
    if os.path.exists(a.out):  # sections that are not re-run keep their last result
        try:
            out = json.load(open(a.out))
        except Exception:
            out = {}
    out.update({"gpu": torch.cuda.get_device_name(0), "peaks": bench.peaks(), "seed": bench.SEED,
                "host_threads": os.cpu_count()})
    only = set(a.only.split(","))
    if "1" in only:
        cfg1(torch, out)
        print("cfg1", json.dumps(out["cfg1"])[:400], file=sys.stderr)
    if "2" in only:
        cfg2(torch, out, a.cfg2_rows)
        for r in out["cfg2"]["rows"]:
            print("cfg2", json.dumps(r), file=sys.stderr)
        torch.cuda.empty_cache()
    if "3" in only:
        cfg3(torch, out, a.cfg3_rows, a.cfg3_new)
        print("cfg3", json.dumps(out["cfg3"]), file=sys.stderr)
        torch.cuda.empty_cache()
    if "5" in only:
        cfg5(torch, out, a.cfg5_rows)
        print("cfg5", json.dumps(out["cfg5"]), file=sys.stderr)
    if "4" in only:
        cfg4(torch, out, a.cfg4_rows, 256)
        print("cfg4", json.dumps(out["cfg4"]), file=sys.stderr)
        torch.cuda.empty_cache()
    if "dedup" in only:
        dedup(torch, out, a.dedup_rows)
        print("dedup", json.dumps(out["dedup"]), file=sys.stderr)
        torch.cuda.empty_cache()
    if "hnsw" in only:
        hnsw(torch, out, a.hnsw_rows, 200)
        print("hnsw", json.dumps(out["hnsw"]), file=sys.stderr)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fp:
        json.dump(out, fp, indent=1)
    print(json.dumps({"wrote": a.out}))


if __name__ == "__main__":
    main()
