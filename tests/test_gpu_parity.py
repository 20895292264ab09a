"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bar: ids identical, scores and distances BIT-identical (every
emitted number is produced by the reference's own operation sequence)."""
import os

import numpy as np
import pytest

from cortex_b200 import GpuVectorIndex, VectorFilter, synth
from oracle.binding import Filter, OracleIndex

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "scan_golden.npz"))


def same_bits(a, b):
    """bit-identical floats; NaN matches NaN (the payload is not part of the value:
    x86 produces 0xFFC00000 for 0/0, the GPU the canonical 0x7FFFFFFF)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def build_pair(corpus, ids=None):
    n, d = corpus.shape
    ids = synth.make_ids(n) if ids is None else ids
    g = GpuVectorIndex(d)
    g.insert_batch(ids, corpus)
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(ids, corpus)
    return g, o, ids


def assert_batch_equal(g, o, Q, k, gflt=None, oflt=None):
    gi, gs, gd, gn = g.search_batch_arrays(Q, k, gflt)
    oi, os_, od, orow, on = o.search_batch(Q, k, oflt)
    assert np.array_equal(gn, on), (gn, on)
    for b in range(Q.shape[0]):
        n = int(on[b])
        assert np.array_equal(gi[b, :n], oi[b, :n]), f"query {b}: ids differ"
        assert same_bits(gs[b, :n], os_[b, :n]), f"query {b}: scores differ"
        assert same_bits(gd[b, :n], od[b, :n]), f"query {b}: distances differ"


@pytest.mark.parametrize("case", ["A", "B", "C", "D"])
def test_golden_vectors_bitwise(case):
    corpus, query = GOLD[f"{case}_corpus"], GOLD[f"{case}_query"]
    n = corpus.shape[0]
    ids = np.stack([np.frombuffer(r.to_bytes(16, "big"), np.uint8) for r in range(n)])
    g = GpuVectorIndex(corpus.shape[1])
    g.insert_batch(ids, corpus)
    qs = query if query.ndim == 2 else query[None, :]
    rows = GOLD[f"{case}_rows"].reshape(len(qs), -1)
    score = GOLD[f"{case}_score"].reshape(len(qs), -1)
    dist = GOLD[f"{case}_dist"].reshape(len(qs), -1)
    for b, q in enumerate(qs):
        res = g.search(q, n)
        assert [int.from_bytes(r.node_id, "big") for r in res] == list(rows[b])
        assert same_bits([r.score for r in res], score[b])
        assert same_bits([r.distance for r in res], dist[b])
        thr = g.search_threshold(q, 0.75)
        exp = [int(r) for r, s in zip(rows[b], score[b]) if s >= np.float32(0.75)]
        assert [int.from_bytes(r.node_id, "big") for r in thr] == exp


def test_config1_10k_384_100_queries_top10():
    """BASELINE.json configs[0]: exact cosine top-10 over 10k x 384, 100 queries."""
    corpus = synth.make_corpus(10_000, 384, zero_row=True)
    Q = synth.make_queries(corpus, 100)
    g, o, _ = build_pair(corpus)
    assert_batch_equal(g, o, Q, 10)
    st = g.stats()
    assert st["queries_stream"] + st["queries_tensor"] > 0, "fast pass did not run"
    assert st["fallbacks"] <= 10, st
    # the exact path gives the same answer
    g.set_option("force_path", 3)
    assert_batch_equal(g, o, Q, 10)


@pytest.mark.parametrize("n,d,b,k", [
    (1000, 384, 1, 10), (4099, 384, 3, 5), (5000, 100, 5, 30), (3000, 1024, 2, 10),
    (2500, 768, 9, 1), (20_000, 384, 8, 100), (300, 64, 4, 10), (257, 384, 1, 50),
    (6000, 1536, 2, 10), (7000, 12, 7, 10),
])
def test_shapes(n, d, b, k):
    corpus = synth.make_corpus(n, d, seed=synth.SEED + n + d)
    Q = synth.make_queries(corpus, b, seed=synth.SEED + n)
    g, o, _ = build_pair(corpus)
    assert_batch_equal(g, o, Q, k)


@pytest.mark.parametrize("n,d,b,k", [
    (5000, 384, 16, 10), (20_000, 384, 200, 10), (3000, 128, 33, 5), (8000, 384, 130, 30),
    (50_000, 384, 1024, 10), (4000, 100, 40, 10), (6000, 640, 17, 10), (700, 384, 129, 3),
    (30_000, 384, 300, 10), (40_000, 256, 640, 100),
    # large embeddings: the query tile no longer fits next to the ring and is streamed with E
    (6000, 1024, 40, 10), (9000, 768, 260, 100), (5000, 1536, 130, 10),
    # 40 query tiles: split into launch groups of 37 + 3 tiles so that no SM idles
    (8000, 128, 5000, 10),
])
@pytest.mark.parametrize("pair", [1, 0])
def test_tensor_pass_parity(n, d, b, k, pair):
    """K2 (tcgen05 bf16 pass) nominates, K3 rescores: results must equal the oracle bit for bit.
    pair=1: two or more query tiles run as CTA pairs (cta_group::2); pair=0: single-CTA form."""
    corpus = synth.make_corpus(n, d, zero_row=True, seed=synth.SEED + 3 * n + d)
    Q = synth.make_queries(corpus, b, seed=synth.SEED + n + 1)
    g, o, _ = build_pair(corpus)
    g.set_option("force_path", 2)
    g.set_option("tensor_pair", pair)
    try:
        assert_batch_equal(g, o, Q, k)
    finally:
        g.set_option("tensor_pair", 0)
    st = g.stats()
    # the tensor pass itself must verify (almost) every query: a retry on the streaming pass would
    # also produce the right answer and hide a wrong tile mapping
    assert st["queries_tensor"] >= 0.9 * b, st
    assert st["queries_exact"] <= 0.05 * b + 1, st


def test_tensor_pass_filters_and_dead_rows():
    corpus = synth.make_corpus(9000, 384, seed=41)
    Q = synth.make_queries(corpus, 64, seed=41)
    g, o, ids = build_pair(corpus)
    for r in range(0, 9000, 3):
        g.set_metadata(ids[r].tobytes(), "fact" if r % 2 else "event", "a1")
        o.set_metadata(ids[r].tobytes(), "fact" if r % 2 else "event", "a1")
    for r in range(5, 9000, 11):
        g.remove(ids[r].tobytes())
        o.remove(ids[r].tobytes())
    g.set_option("force_path", 2)
    assert_batch_equal(g, o, Q, 10, VectorFilter().with_kinds(["fact"]), Filter(kinds=["fact"]))
    assert_batch_equal(g, o, Q, 10)
    assert g.stats()["queries_tensor"] > 0


def test_non_normalised_and_duplicates():
    corpus = synth.make_corpus(8000, 384, normalise=False, dup_frac=0.2, seed=7)
    Q = synth.make_queries(corpus, 6, seed=7) * np.float32(3.0)
    g, o, _ = build_pair(corpus)
    assert_batch_equal(g, o, Q, 20)


def test_all_rows_identical_ties_in_row_order():
    row = synth.make_corpus(1, 384, dup_frac=0.0)[0]
    corpus = np.tile(row, (2000, 1))
    g, o, ids = build_pair(corpus)
    res = g.search(row, 10)
    assert [r.node_id for r in res] == [ids[i].tobytes() for i in range(10)]
    assert_batch_equal(g, o, row[None, :], 10)


def test_opposite_and_orthogonal_queries_score_zero_ties():
    """score clamps at 0 (index.rs:255): ties resolve in row order, identically to the oracle."""
    corpus = synth.make_corpus(3000, 384, seed=11)
    g, o, _ = build_pair(corpus)
    q = -corpus[5]
    assert_batch_equal(g, o, q[None, :], 10)


def test_filters_fast_and_exact():
    corpus = synth.make_corpus(6000, 384, seed=3)
    Q = synth.make_queries(corpus, 4, seed=3)
    g, o, ids = build_pair(corpus)
    kinds = ["fact", "decision", "event"]
    for r in range(0, 6000, 2):  # half the rows carry metadata
        kd, ag = kinds[r % 3], f"agent-{r % 5}"
        g.set_metadata(ids[r].tobytes(), kd, ag)
        o.set_metadata(ids[r].tobytes(), kd, ag)
    top = g.search(Q[0], 3)
    excl = [r.node_id for r in top]
    cases = [
        (VectorFilter().with_kinds(["decision"]), Filter(kinds=["decision"])),
        (VectorFilter().excluding(excl), Filter(exclude=excl)),
        (VectorFilter().with_source_agent("agent-2"), Filter(source_agent="agent-2")),
        (VectorFilter().with_kinds(["fact", "event"]).with_source_agent("agent-1").excluding(excl),
         Filter(kinds=["fact", "event"], source_agent="agent-1", exclude=excl)),
        (VectorFilter().with_kinds(["unknown-kind"]), Filter(kinds=["unknown-kind"])),
    ]
    for mode in (0, 3):
        g.set_option("force_path", mode)
        for gf, of in cases:
            assert_batch_equal(g, o, Q, 10, gf, of)


def test_remove_rebuild_overwrite():
    corpus = synth.make_corpus(5000, 384, seed=5)
    Q = synth.make_queries(corpus, 5, seed=5)
    g, o, ids = build_pair(corpus)
    for r in range(0, 5000, 7):
        g.remove(ids[r].tobytes())
        o.remove(ids[r].tobytes())
    assert len(g) == len(o)
    assert_batch_equal(g, o, Q, 10)
    # overwrite some live rows in place (HashMap::insert semantics)
    for r in (1, 2, 3):
        g.insert(ids[r].tobytes(), corpus[r + 10])
        o.insert(ids[r].tobytes(), corpus[r + 10])
    assert_batch_equal(g, o, Q, 10)
    g.rebuild()
    assert len(g) == len(o)
    assert_batch_equal(g, o, Q, 10)
    # re-insert a removed id: appended as a new row in both
    g.insert(ids[0].tobytes(), corpus[0])
    o.insert(ids[0].tobytes(), corpus[0])
    assert_batch_equal(g, o, Q, 10)


def test_threshold_search_matches_oracle():
    corpus = synth.make_corpus(4000, 384, zero_row=True, seed=9)
    Q = synth.make_queries(corpus, 4, seed=9)
    g, o, _ = build_pair(corpus)
    for q in Q:
        for t in (0.92, 0.75, 0.5, 0.0, -1.0, 1.5):
            res = g.search_threshold(q, t)
            exp = o.search_threshold(q, t)
            assert [r.node_id for r in res] == [i.tobytes() for i in exp.ids]
            assert same_bits([r.score for r in res], exp.score)


def test_query_length_mismatch_follows_zip_truncation():
    corpus = synth.make_corpus(500, 16, seed=13)
    g, o, _ = build_pair(corpus)
    for qlen in (5, 16, 23):
        q = np.linspace(-1, 1, qlen).astype(np.float32)
        res = g.search(q, 7)
        exp = o.search(q, 7)
        assert [r.node_id for r in res] == [i.tobytes() for i in exp.ids]
        assert same_bits([r.distance for r in res], exp.distance)


def test_save_load_roundtrip_and_layout(tmp_path):
    corpus = synth.make_corpus(700, 96, seed=17)
    g, o, ids = build_pair(corpus)
    for r in range(0, 700, 3):
        g.set_metadata(ids[r].tobytes(), "fact", "a")
        o.set_metadata(ids[r].tobytes(), "fact", "a")
    g.remove(ids[4].tobytes())
    o.remove(ids[4].tobytes())
    pg, po = str(tmp_path / "g.bin"), str(tmp_path / "o.bin")
    g.save(pg)
    o.save(po)
    assert open(pg, "rb").read() == open(po, "rb").read()  # same bincode bytes, same row order
    g2 = GpuVectorIndex.load(pg)
    assert len(g2) == len(o)
    Q = synth.make_queries(corpus, 3, seed=17)
    assert_batch_equal(g2, o, Q, 10, VectorFilter().with_kinds(["fact"]), Filter(kinds=["fact"]))


def test_large_properties_200k():
    """Size-independent properties + spot parity at a size the oracle handles per query."""
    n, d = 200_000, 384
    corpus = synth.make_corpus(n, d, seed=21)
    g, o, ids = build_pair(corpus)
    Q = np.concatenate([corpus[[0, 77_777, n - 1]], synth.make_queries(corpus, 13, seed=21)])
    gi, gs, gd, gn = g.search_batch_arrays(Q, 10)
    assert np.all(gn == 10)
    assert np.all(np.diff(gs, axis=1) <= 0), "scores must be non-increasing"
    for b, r in enumerate([0, 77_777, n - 1]):  # a stored row finds itself (or an exact duplicate) first
        assert gs[b, 0] >= np.float32(0.9999)
    gi2, gs2, _, _ = g.search_batch_arrays(Q, 10)
    assert np.array_equal(gi, gi2) and np.array_equal(gs, gs2)  # idempotent
    assert_batch_equal(g, o, Q[:6], 10)
    st = g.stats()
    assert st["fallbacks"] <= 2, st
