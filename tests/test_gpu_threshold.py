"""GPU parity for threshold scans: search_threshold (vector/index.rs:376-388), its batch form
and the dedup self-join (linker/dedup.rs:65-127), against the CPU oracle.  The fast path
nominates every row within the pass's error bound of the threshold, rescores ALL nominees with
reference arithmetic and applies the exact `score >= threshold`; ids must be identical and
scores / distances bit-identical."""
import numpy as np
import pytest

from cortex_b200 import GpuVectorIndex, VectorFilter, synth
from oracle.binding import Filter, OracleIndex

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def build_pair(corpus):
    n, d = corpus.shape
    ids = synth.make_ids(n)
    g = GpuVectorIndex(d)
    g.insert_batch(ids, corpus)
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(ids, corpus)
    return g, o, ids


def check_batch(g, o, Q, thr, cap, gflt=None, oflt=None):
    ids, sc, di, n, total = g.search_threshold_batch_arrays(Q, thr, cap, gflt)
    for b in range(Q.shape[0]):
        exp = o.search_threshold(Q[b], thr, oflt)
        assert int(total[b]) == len(exp.ids), (b, int(total[b]), len(exp.ids))
        m = min(cap, len(exp.ids))
        assert int(n[b]) == m
        assert np.array_equal(ids[b, :m], exp.ids[:m]), f"query {b}: ids differ"
        assert same_bits(sc[b, :m], exp.score[:m]), f"query {b}: scores differ"
        assert same_bits(di[b, :m], exp.distance[:m]), f"query {b}: distances differ"


@pytest.mark.parametrize("n,d,b", [(4000, 384, 1), (9000, 384, 4), (6000, 128, 7), (5000, 768, 3)])
def test_threshold_streaming_pass(n, d, b):
    """Small batches: K1 in threshold mode nominates (B < 3), the rescoring kernel decides."""
    corpus = synth.make_corpus(n, d, zero_row=True, seed=21 + n)
    Q = synth.make_queries(corpus, b, seed=n)
    g, o, _ = build_pair(corpus)
    for thr in (0.92, 0.75, 0.5, 0.3):
        st0 = g.stats()
        check_batch(g, o, Q, thr, 512)
        st1 = g.stats()
        key = "queries_stream" if b < 3 else "queries_tensor"  # tensor_min_batch = 3
        assert st1[key] - st0[key] == b, (thr, st0, st1)  # the fast path served them
        assert st1["queries_exact"] == st0["queries_exact"]
    # the streaming pass in threshold mode with up to eight queries per pass (what indexes without a shadow use)
    g.set_option("force_path", 1)
    st0 = g.stats()
    check_batch(g, o, Q, 0.75, 512)
    st1 = g.stats()
    assert st1["queries_stream"] - st0["queries_stream"] == b and st1["queries_exact"] == st0["queries_exact"]


@pytest.mark.parametrize("n,d,b", [(20_000, 384, 64), (30_000, 384, 300), (8000, 256, 129)])
def test_threshold_tensor_pass(n, d, b):
    """B >= 3: the tcgen05 pass with a fixed cut-off nominates."""
    corpus = synth.make_corpus(n, d, zero_row=True, seed=5 + n)
    Q = synth.make_queries(corpus, b, seed=n + 2)
    g, o, _ = build_pair(corpus)
    for thr in (0.92, 0.75):
        st0 = g.stats()
        check_batch(g, o, Q, thr, 256)
        st1 = g.stats()
        assert st1["queries_tensor"] - st0["queries_tensor"] == b, (thr, st0, st1)
        assert st1["queries_exact"] == st0["queries_exact"]


def test_threshold_low_values_and_overflow_take_the_exact_path():
    corpus = synth.make_corpus(5000, 64, seed=77)
    corpus[1000:4000] = corpus[7]  # 3000 copies of one row: more qualifying rows than the fast path orders
    Q = np.stack([corpus[7], corpus[11]])
    g, o, _ = build_pair(corpus)
    for thr in (0.9, 0.1, 0.0, -0.5, float("nan")):
        check_batch(g, o, Q, thr, 5000)
    assert g.stats()["queries_exact"] > 0


def test_threshold_filters_and_removed_rows():
    corpus = synth.make_corpus(7000, 384, seed=3)
    Q = synth.make_queries(corpus, 6, seed=4)
    g, o, ids = build_pair(corpus)
    for r in range(0, 7000, 2):
        g.set_metadata(ids[r].tobytes(), "fact" if r % 4 else "event", "a1")
        o.set_metadata(ids[r].tobytes(), "fact" if r % 4 else "event", "a1")
    for r in range(3, 7000, 7):
        g.remove(ids[r].tobytes())
        o.remove(ids[r].tobytes())
    check_batch(g, o, Q, 0.75, 300, VectorFilter().with_kinds(["fact"]), Filter(kinds=["fact"]))
    check_batch(g, o, Q[:3], 0.6, 300)


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


@pytest.mark.parametrize("n,d", [(3000, 64), (5000, 384)])
def test_dedup_scan_matches_reference_loop(n, d):
    corpus = synth.make_corpus(n, d, seed=1234 + n)
    rng = np.random.default_rng(5)
    near = rng.integers(0, n, 200)
    corpus[near] = corpus[(near * 7 + 1) % n] + rng.normal(0, 0.004, (200, d)).astype(np.float32)  # near duplicates
    g, o, ids = build_pair(corpus)
    dead = set(range(17, n, 97))
    for r in dead:
        g.remove(ids[r].tobytes())
        o.remove(ids[r].tobytes())
    a, b, sc, total = g.dedup_scan(0.92, per_node_cap=256)
    exp = oracle_dedup(o, corpus, ids, 0.92, dead)
    assert total == len(exp) and len(a) == len(exp)
    got = [(a[i].tobytes(), b[i].tobytes(), sc[i]) for i in range(len(a))]
    # same pairs, same scores; the reference's order is (a asc, score desc), ours adds row order for ties
    assert sorted((x[0], x[1]) for x in got) == sorted((x[0], x[1]) for x in exp)
    gm = {(x[0], x[1]): x[2] for x in got}
    for x in exp:
        assert np.float32(gm[(x[0], x[1])]).view(np.uint32) == np.float32(x[2]).view(np.uint32)
    assert [x[0] for x in got] == [x[0] for x in exp]  # grouped by the first-visited node, in order
    assert len(exp) > 50
