#!/usr/bin/env python
"""bench.py -- the similarity-scan hot path on N B200s (driver contract in DESIGN.md §7).

Workload (BASELINE.json configs[1]): search over a 1,000,000 x 384 fp32 corpus per
GPU, query batch B, top-10.  A "step" is one search_batch of B queries.

  value : queries/s with queries and results resident in HBM (cx_search_batch_device)
  e2e   : the same metric through the reference-facing host call (cx_search_batch:
          host query buffer in, host ids/scores out, copies inside the timed region)
  roofline : the scan-pass kernel, algorithmic bytes (rows*ld*4 per launch) over its
          CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the CPU restatement of the reference's exact
          scan (oracle/, per-pair row clone + full stable sort like index.rs:259-294)
          on the host cores, bounded to one query per thread per step.

N > 1 (torchrun): the corpus is row-sharded, every rank scans its own 1M-row shard
for the same B queries, local top-k lists are exchanged with one NCCL all_gather and
merged (cortex_b200/sharded.py).  Weak scaling: per-GPU work is fixed; `value`
counts per-shard query scans (B x n_gpus per step), which equals queries/s at N=1.
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
    ap.add_argument("--autolink-new", type=int, default=16384,
                    help="new nodes per auto-link cycle in the extra 'autolink' measurement (0 = skip)")
    ap.add_argument("--pipeline-depth", type=int, default=2,
                    help="searches in flight in the device-resident measurement (1 = strictly one after the other)")
    ap.add_argument("--pin", action="store_true",
                    help="N > 1: give every rank its own contiguous block of the host's CPUs (sched_setaffinity)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=INT",
                    help="cx_set_option before the run (A/B measurements), e.g. --opt tensor_pair=0")
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


# ----------------------------------------------------------------------------------
# synthetic data (clustered unit-norm rows), generated with torch so
# that 1M x 384 takes a fraction of a second on the GPU and a few seconds on CPU.
def make_corpus_torch(n, d, seed, device):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n_clusters = max(1, n // 64)
    cent = torch.randn((n_clusters, d), generator=g, device=device, dtype=torch.float32)
    cent /= cent.norm(dim=1, keepdim=True)
    out = torch.empty((n, d), device=device, dtype=torch.float32)
    chunk = 1 << 18
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        which = torch.randint(0, n_clusters, (e - s,), generator=g, device=device)
        scale = 0.035 * (0.3 + 1.3 * torch.rand((e - s, 1), generator=g, device=device))
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
def cpu_reference_leg(corpus_np, queries_np, k, steps, warmup, n_queries):
    """The reference's exact scan restated on the CPU (oracle/), all host threads,
    one bounded batch per step.  Returns (queries/s, cores, ms_per_step, sample)."""
    from oracle.binding import OracleIndex, max_threads

    cores = max_threads()
    nq = n_queries or cores
    nq = min(nq, queries_np.shape[0])
    ix = OracleIndex(corpus_np.shape[1], faithful_copy=True)
    ids = np.zeros((corpus_np.shape[0], 16), np.uint8)
    ids[:, 8:] = np.arange(corpus_np.shape[0], dtype=np.uint64).astype(">u8").view(np.uint8).reshape(-1, 8)
    ix.insert_batch(ids, corpus_np)
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
    sample = (f"{nq} queries/step x {len(times)} steps against the full {corpus_np.shape[0]}x{corpus_np.shape[1]} "
              f"corpus, OpenMP over queries (= rayon par_iter, index.rs:397), per-pair row clone + full stable sort")
    return qps, cores, 1e3 * tot / len(times), sample


def cpu_optimised_leg(corpus_np, queries_np, k, n_queries=64):
    """Courtesy number, NOT the reference: oracle/cpu_fast.c (hoisted norms, one corpus pass per 16
    queries, AVX-512 reassociated fp32, k-heaps, OpenMP) on the same corpus.  Never raises."""
    try:
        from oracle.binding import CpuFastScan, max_threads

        f = CpuFastScan(corpus_np)
        q = queries_np[:n_queries]
        f.search_batch(q[:16], k)
        t0 = time.perf_counter()
        f.search_batch(q, k)
        dt = time.perf_counter() - t0
        return {"value": q.shape[0] / dt, "unit": "queries/s", "cores": max_threads(), "isa": "x86-64-" + f.isa,
                "sample": f"{q.shape[0]} queries against the full corpus, top-{k}",
                "note": "optimised CPU scan (oracle/cpu_fast.c), approximate scores, NOT the reference's algorithm"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:200]}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library
    chatter) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the run
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"search: {a.rows}x{a.dim}-d fp32 corpus per GPU, query batch {a.batch}, top-{a.k} "
                f"(BASELINE.json configs[1])")

    if a.impl == "reference":
        if rank != 0:
            return
        import torch

        corpus = make_corpus_torch(a.rows, a.dim, SEED, "cpu")
        queries = make_queries_torch(corpus, max(256, a.batch), SEED).numpy()
        qps, cores, ms, sample = cpu_reference_leg(corpus.numpy(), queries, a.k, a.steps, a.warmup, a.cpu_queries)
        emit({
            "impl": "reference", "metric": "queries/s", "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "seed": SEED},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return

    if a.pin and world > 1 and hasattr(os, "sched_setaffinity"):
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // world)
        os.sched_setaffinity(0, set(cpus[local_rank * per:(local_rank + 1) * per]))

    import torch
    import torch.distributed as dist

    from cortex_b200 import GpuVectorIndex
    from cortex_b200.sharded import ShardedSearch

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- data: this rank's shard + the (shared) query batch ---------------------
    corpus = make_corpus_torch(a.rows, a.dim, SEED + 7919 * rank, dev)
    q_all = make_queries_torch(corpus, max(256, a.batch), SEED)
    if world > 1:
        dist.broadcast(q_all, src=0)
    d_q = q_all[:a.batch].contiguous()
    # new nodes of the auto-link measurement: perturbed copies of rows of rank 0's shard
    q_al = make_queries_torch(corpus, a.autolink_new, SEED + 5) if a.autolink_new > 0 else None
    corpus_np = corpus.cpu().numpy()
    ids = np.zeros((a.rows, 16), np.uint8)
    ids[:, 8:] = (np.arange(a.rows, dtype=np.uint64) + rank * a.rows).astype(">u8").view(np.uint8).reshape(-1, 8)
    ix = GpuVectorIndex(a.dim, device=local_rank)
    ix.reserve(a.rows)
    ix.insert_batch(ids, corpus_np)
    del corpus
    torch.cuda.empty_cache()
    ix.set_option("profile", 1)
    for kv in a.opt:
        key, val = kv.split("=")
        ix.set_option(key, int(val))

    stream = torch.cuda.current_stream()
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

    def step_device():
        return sh.search(d_q, a.k)

    def run_steps(n):
        """n steps with up to --pipeline-depth searches in flight: the next batch is enqueued before the
        host waits for the previous one, so launch and wait latency are off the GPU's critical path.  Every
        step's result is complete and verified (search_end) inside the timed region."""
        if a.pipeline_depth <= 1:
            for _ in range(n):
                step_device()
            return
        pend = []
        for i in range(n):
            pend.append(sh.search_begin(d_q, a.k, slot=i % a.pipeline_depth))
            if len(pend) >= a.pipeline_depth:
                sh.search_end(pend.pop(0))
        while pend:
            sh.search_end(pend.pop(0))

    h_q = torch.empty((a.batch, a.dim), dtype=torch.float32).pin_memory()
    h_q.copy_(d_q)
    h_q_np = h_q.numpy()

    h_out = (torch.empty((a.batch, a.k), dtype=torch.int64).pin_memory(),
             torch.empty((a.batch, a.k), dtype=torch.float32).pin_memory(),
             torch.empty((a.batch, a.k), dtype=torch.float32).pin_memory(),
             torch.empty((a.batch,), dtype=torch.int32).pin_memory())

    def step_e2e():
        if world == 1:
            return ix.search_batch_arrays(h_q_np, a.k)  # host in, host out: the reference-facing call
        dq = h_q.to(dev, non_blocking=True)
        res = sh.search(dq, a.k)
        for hb, t in zip(h_out, res):          # pinned host buffers, one wait for all four copies
            hb.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident ---------------------------------------------------
    clocks = ClockSampler(world)
    if rank == 0:
        clocks.start()
    for _ in range(max(3, a.warmup)):
        step_device()
    # untimed soak (~1 s) so that clocks settle and nvidia-smi gets samples under load
    t_soak = time.perf_counter()
    n_soak = torch.zeros(1, device=dev)
    while True:
        for _ in range(8):
            step_device()
        n_soak.fill_(1.0 if time.perf_counter() - t_soak < 1.0 else 0.0)
        if world > 1:
            dist.broadcast(n_soak, src=0)
        if n_soak.item() == 0.0:
            break
    run_steps(2 * max(1, a.pipeline_depth))  # warm the pipelined form too (second workspace, exchange buffers)
    barrier()
    st0 = ix.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(a.steps)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    st1 = ix.stats()
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / a.steps
    value = a.batch * world * a.steps / (ms_total * 1e-3)

    # ---- N > 1: the same steps without the exchange (every rank scans, nobody gathers) -- tells the
    # all_gather + merge apart from the local scan when reading the scaling numbers
    local_only_ms = None
    if world > 1:
        for _ in range(3):
            local_search(d_q, a.k)
        barrier()
        sl0 = ix.stats()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(a.steps):
            local_search(d_q, a.k)
        l1.record()
        barrier()
        sl1 = ix.stats()
        k_us = (sl1["pass_kernel_ns"] - sl0["pass_kernel_ns"]) * 1e-3 / max(1, sl1["pass_kernel_launches"] - sl0["pass_kernel_launches"])
        tl = torch.tensor([l0.elapsed_time(l1) / a.steps, k_us], device=dev, dtype=torch.float64)
        tl_all = torch.empty((world, 2), device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(tl_all.view(-1), tl)
        local_only_ms = {"step_ms_per_rank": [round(float(x), 4) for x in tl_all[:, 0].tolist()],
                         "scan_kernel_us_per_rank": [round(float(x), 1) for x in tl_all[:, 1].tolist()]}

    # ---- e2e: host buffers in and out ---------------------------------------------
    for _ in range(max(3, a.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = a.batch * world * a.steps / float(te.item())
    clk = clocks.stop() if rank == 0 else None
    h2d = a.batch * a.dim * 4
    d2h = a.batch * a.k * (16 + 4 + 4) + a.batch * 4

    # ---- small-batch probe (B=1 interactive search, the HBM-bound streaming pass) ----
    small = None
    if a.batch != 1 and not a.no_small_probe:
        q1 = d_q[:1].contiguous()
        o1 = None
        for _ in range(5):
            o1 = ix.search_batch_device(q1, a.k, stream=stream.cuda_stream, out=o1)
        torch.cuda.synchronize()
        s0 = ix.stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n1 = 100
        ev0.record()
        for _ in range(n1):
            o1 = ix.search_batch_device(q1, a.k, stream=stream.cuda_stream, out=o1)
        ev1.record()
        torch.cuda.synchronize()
        s1 = ix.stats()
        pk1 = peaks()
        l1 = s1["pass_kernel_launches"] - s0["pass_kernel_launches"]
        ns1 = s1["pass_kernel_ns"] - s0["pass_kernel_ns"]
        if l1 and ns1:
            bytes1 = a.rows * ((a.dim + 3) // 4 * 4) * 4
            gbs = bytes1 / (ns1 * 1e-9 / l1) / 1e9
            small = {"batch": 1, "queries_per_s": n1 / (ev0.elapsed_time(ev1) * 1e-3),
                     "ms_per_query": ev0.elapsed_time(ev1) / n1,
                     "roofline": {"bound": "hbm", "kernel": "stream_scan_kernel", "achieved": gbs,
                                  "peak": pk1["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk1["hbm_gbs"],
                                  "frac_of_nominal_8TBs": gbs / 8000.0, "us_per_launch": ns1 * 1e-3 / l1,
                                  "traffic": ncu_traffic("stream_scan_kernel", 1),
                                  "algorithmic_bytes_per_launch": bytes1, "peak_source": pk1["source"]}}

    # ---- auto-link cycle (BASELINE.json metric "auto-link pairs/s at 1/2/4/8 B200", configs[2] shape at a
    # bounded size): every new node searches the row-sharded corpus for its 100 nearest neighbours
    # (linker/auto_linker.rs:215-222), local lists are merged with one all_gather, candidates with
    # score >= 0.75 become links (linker/rules.rs:50).  Unit = scored (new node, corpus row) pair.
    autolink = None
    if a.autolink_new > 0:
        nq = a.autolink_new
        if world > 1:
            dist.broadcast(q_al, src=0)
        q_al = q_al.contiguous()
        out_al = None

        def local_al(q, k):
            nonlocal out_al
            out_al = ix.search_batch_device(q, k, stream=stream.cuda_stream, out=out_al)
            return out_al

        def local_al_begin(q, k, slot=0):
            nonlocal out_al
            out_al, ticket = ix.search_batch_device_begin(q, k, stream=stream.cuda_stream, out=out_al)
            return out_al, ticket

        sh_al = ShardedSearch(local_al, row_offset=rank * a.rows, local_begin=local_al_begin,
                              local_end=ix.search_batch_device_end, ticket_ok_ptr=ix.ticket_ok_ptr)
        res = None
        for _ in range(2):
            res = sh_al.autolink(q_al, None, 100, 0.75, 50)
        barrier()
        sa0 = ix.stats()
        ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_cyc = 3
        ea0.record()
        for _ in range(n_cyc):
            res = sh_al.autolink(q_al, None, 100, 0.75, 50)  # sharded search(k=100) + threshold + cap 50
            links = res[2].sum()
        ea1.record()
        barrier()
        sa1 = ix.stats()
        tt = torch.tensor([ea0.elapsed_time(ea1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_cyc = float(tt.item()) / n_cyc
        pairs = float(nq) * a.rows * world
        pk_a = peaks()
        tf = 2.0 * a.dim * float(nq) * a.rows / (ms_cyc * 1e-3) / 1e12  # per GPU
        autolink = {"metric": "scored pairs/s", "value": pairs / (ms_cyc * 1e-3), "unit": "pairs/s",
                    "new_nodes_per_cycle": nq, "rows_per_gpu": a.rows, "k": 100, "threshold": 0.75,
                    "ms_per_cycle": ms_cyc, "link_candidates": int(links.item()),
                    "tflops_per_gpu_whole_cycle": tf, "frac_of_sustained_bf16_whole_cycle": tf / pk_a["bf16_tflops_sustained"],
                    "fallbacks": sa1["fallbacks"] - sa0["fallbacks"]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the scan pass) --------------------------
    pk = peaks()
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
                "kernel_share_of_step": (ns * 1e-6) / ms_total}
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
                "launch_note": "one launch = one scan of the shard: its phase launches of tensor_scan_kernel plus "
                               "the tau_refine kernels between them (CUDA events bracket all of it)",
                "hbm_floor_us": hbm_bytes / (pk["hbm_gbs"] * 1e9) * 1e6,
                "hbm_gbs_of_shadow_stream": hbm_bytes / sec / 1e9,
                "kernel_share_of_step": (ns * 1e-6) / ms_total}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        qps, cores, ms, sample = cpu_reference_leg(corpus_np, q_all.cpu().numpy(), a.k, 2, 1, a.cpu_queries)
        cpu = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
               "optimised_courtesy": cpu_optimised_leg(corpus_np, q_all.cpu().numpy(), a.k)}

    line = {
        "metric": "queries/s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "rows_per_gpu": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k,
                   "seed": SEED, "parallelism": f"row-shard x{world}",
                   "cache": "inputs larger than L2: the 1.5 GB corpus shard is streamed from HBM every step",
                   "value_counts": "per-shard query scans (batch x n_gpus per step)",
                   "pipeline_depth": a.pipeline_depth},
        "roofline": roof, "small_batch": small, "autolink": autolink, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": st1["kernel_launches"] - st0["kernel_launches"],
        "local_scan_only_ms_per_step": local_only_ms, "host_cpus": os.cpu_count(),
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
