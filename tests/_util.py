"""Helpers shared by the GPU parity tests (the checker side is the CPU oracle)."""
import numpy as np

from cortex_b200 import GpuVectorIndex, synth
from oracle.binding import OracleIndex


def same_bits(a, b):
    """bit-identical floats; NaN matches NaN (the payload is not part of the value:
    x86 produces 0xFFC00000 for 0/0, the GPU the canonical 0x7FFFFFFF)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def build_pair(corpus, ids=None, devices=None):
    n, d = corpus.shape
    ids = synth.make_ids(n) if ids is None else ids
    g = GpuVectorIndex(d, devices=devices) if devices is not None else GpuVectorIndex(d)
    g.insert_batch(ids, corpus)
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(ids, corpus)
    return g, o, ids


def assert_batch_equal(g, o, Q, k, gflt=None, oflt=None, sample=None):
    """search_batch on the GPU path vs the oracle; `sample` = indices of the queries the oracle checks
    (all of them by default)."""
    gi, gs, gd, gn = g.search_batch_arrays(Q, k, gflt)
    idx = np.arange(Q.shape[0]) if sample is None else np.asarray(sample)
    oi, os_, od, _, on = o.search_batch(Q[idx], k, oflt)
    assert np.array_equal(gn[idx], on), (gn[idx], on)
    for j, b in enumerate(idx):
        n = int(on[j])
        assert np.array_equal(gi[b, :n], oi[j, :n]), f"query {b}: ids differ"
        assert same_bits(gs[b, :n], os_[j, :n]), f"query {b}: scores differ"
        assert same_bits(gd[b, :n], od[j, :n]), f"query {b}: distances differ"


def check_threshold_batch(g, o, Q, thr, cap, gflt=None, oflt=None):
    ids, sc, di, n, total = g.search_threshold_batch_arrays(Q, thr, cap, gflt)
    for b in range(Q.shape[0]):
        exp = o.search_threshold(Q[b], thr, oflt)
        assert int(total[b]) == len(exp.ids), (b, int(total[b]), len(exp.ids))
        m = min(cap, len(exp.ids))
        assert int(n[b]) == m
        assert np.array_equal(ids[b, :m], exp.ids[:m]), f"query {b}: ids differ"
        assert same_bits(sc[b, :m], exp.score[:m]), f"query {b}: scores differ"
        assert same_bits(di[b, :m], exp.distance[:m]), f"query {b}: distances differ"


def oracle_dedup(o, corpus, ids, thr, dead=()):
    """linker/dedup.rs:65-127 over nodes in insertion order: search_threshold per node, skip self,
    report each unordered pair once (from the node visited first)."""
    seen, out = set(), []
    for r in range(corpus.shape[0]):
        if r in dead:
            continue
        hits = o.search_threshold(corpus[r], thr)
        for hid, hs, hr in zip(hits.ids, hits.score, hits.rows):
            if int(hr) == r:
                continue
            key = (min(r, int(hr)), max(r, int(hr)))
            if key in seen:
                continue
            seen.add(key)
            out.append((ids[r].tobytes(), hid.tobytes(), np.float32(hs)))
    return out


def assert_dedup_equal(g, o, corpus, ids, thr, dead=(), per_node_cap=256, exact_order=False):
    a, b, sc, total = g.dedup_scan(thr, per_node_cap=per_node_cap)
    exp = oracle_dedup(o, corpus, ids, thr, dead)
    assert total == len(exp) and len(a) == len(exp), (total, len(a), len(exp))
    got = [(a[i].tobytes(), b[i].tobytes(), sc[i]) for i in range(len(a))]
    # same pairs, same scores; the reference's order is (a asc, score desc), ours adds row order for ties
    assert sorted((x[0], x[1]) for x in got) == sorted((x[0], x[1]) for x in exp)
    gm = {(x[0], x[1]): x[2] for x in got}
    for x in exp:
        assert np.float32(gm[(x[0], x[1])]).view(np.uint32) == np.float32(x[2]).view(np.uint32)
    assert [x[0] for x in got] == [x[0] for x in exp]  # grouped by the first-visited node, in order
    if exact_order:
        assert [(x[0], x[1]) for x in got] == [(x[0], x[1]) for x in exp]
    return exp
