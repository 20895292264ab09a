"""The streaming pass over the bf16 shadow (cx_stream.cu, HALF variant): batches of up to four queries read the
normalised bf16 copy of the rows (half the bytes) to NOMINATE; every emitted number still comes from the exact
fp32 rescoring, so results stay bit-identical to the oracle (vector/index.rs:169-179, 253-294).  Each case is
also run with the option off (fp32 rows) and must give the same bits."""
import numpy as np
import pytest

from cortex_b200 import GpuVectorIndex, VectorFilter, synth
from oracle.binding import Filter, OracleIndex

from _util import assert_batch_equal, same_bits

pytestmark = pytest.mark.gpu


def build_pair(corpus, ids=None):
    n, d = corpus.shape
    ids = synth.make_ids(n) if ids is None else ids
    g = GpuVectorIndex(d)
    g.insert_batch(ids, corpus)
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(ids, corpus)
    return g, o, ids


def both_ways(g, o, Q, k, gflt=None, oflt=None, expect_half=True):
    """default (bf16 shadow) and stream_bf16=0 (fp32 rows): same bits, both equal to the oracle"""
    s0 = g.stats()
    assert_batch_equal(g, o, Q, k, gflt, oflt)
    a = g.search_batch_arrays(Q, k, gflt)
    s1 = g.stats()
    if expect_half:
        assert s1["queries_stream_bf16"] > s0["queries_stream_bf16"], (s0, s1)
    g.set_option("stream_bf16", 0)
    try:
        assert_batch_equal(g, o, Q, k, gflt, oflt)
        b = g.search_batch_arrays(Q, k, gflt)
        s2 = g.stats()
        assert s2["queries_stream_bf16"] == s1["queries_stream_bf16"]
    finally:
        g.set_option("stream_bf16", 1)
    assert np.array_equal(a[0], b[0]) and same_bits(a[1], b[1]) and same_bits(a[2], b[2]) and np.array_equal(a[3], b[3])


@pytest.mark.parametrize("b", [1, 2, 3, 4])
def test_batches_of_one_to_four(b):
    corpus = synth.make_corpus(30_000, 384, dup_frac=0.05, seed=21)
    Q = synth.make_queries(corpus, b, seed=22 + b)
    g, o, _ = build_pair(corpus)
    g.set_option("tensor_min_batch", 5)  # by default three and four queries go to the tensor pass (below)
    both_ways(g, o, Q, 10)
    st = g.stats()
    assert st["queries_tensor"] == 0 and st["queries_exact"] == 0, st


def test_default_routing_of_small_batches():
    """B = 1, 2: streaming pass over the shadow; B >= 3: tensor pass; k beyond the tensor pass's keep lists
    (k > 112): the streaming pass again, four queries per pass"""
    corpus = synth.make_corpus(20_000, 384, seed=33)
    g, o, _ = build_pair(corpus)
    small = synth.make_corpus(3000, 384, seed=34)   # few enough producer groups for keep lists of 120+
    gs, os_, _ = build_pair(small)
    for gg, oo, c, b, k, key in ((g, o, corpus, 1, 10, "queries_stream_bf16"), (g, o, corpus, 2, 10, "queries_stream_bf16"),
                                 (g, o, corpus, 3, 10, "queries_tensor"), (g, o, corpus, 4, 10, "queries_tensor"),
                                 (gs, os_, small, 4, 120, "queries_stream_bf16")):
        Q = synth.make_queries(c, b, seed=40 + b)
        s0 = gg.stats()
        assert_batch_equal(gg, oo, Q, k)
        s1 = gg.stats()
        assert s1[key] - s0[key] == b, (b, k, key, s0, s1)


@pytest.mark.parametrize("n,d,k", [
    (5000, 12, 5), (5000, 100, 10), (6000, 200, 1), (4000, 768, 10), (4000, 1024, 100), (3000, 1536, 20),
    (257, 384, 10), (20_000, 384, 100), (1031, 64, 64),
])
def test_shapes(n, d, k):
    """row lengths with one, two and three slices per stage, padded tails (ld16 > dim), k up to the keep limit"""
    corpus = synth.make_corpus(n, d, seed=n + d)
    Q = synth.make_queries(corpus, 3, seed=k)
    g, o, _ = build_pair(corpus)
    g.set_option("tensor_min_batch", 5)
    both_ways(g, o, Q, k)


def test_non_normalised_rows_and_scaled_queries():
    corpus = synth.make_corpus(8000, 384, normalise=False, dup_frac=0.2, seed=7)
    Q = synth.make_queries(corpus, 4, seed=7) * np.float32(3.0)
    g, o, _ = build_pair(corpus)
    g.set_option("tensor_min_batch", 5)
    both_ways(g, o, Q, 20)


def test_filters_and_removed_rows():
    corpus = synth.make_corpus(9000, 384, seed=41)
    Q = synth.make_queries(corpus, 2, seed=41)
    g, o, ids = build_pair(corpus)
    for r in range(0, 9000, 3):
        g.set_metadata(ids[r].tobytes(), "fact" if r % 2 else "event", "a1")
        o.set_metadata(ids[r].tobytes(), "fact" if r % 2 else "event", "a1")
    for r in range(5, 9000, 11):
        g.remove(ids[r].tobytes())
        o.remove(ids[r].tobytes())
    both_ways(g, o, Q, 10, VectorFilter().with_kinds(["fact"]), Filter(kinds=["fact"]))
    both_ways(g, o, Q, 10)
    top = [r.node_id for r in g.search(Q[0], 3)]
    both_ways(g, o, Q, 10, VectorFilter().excluding(top), Filter(exclude=top))


def test_identical_rows_and_near_ties_fall_back_to_tighter_passes():
    """2 000 copies of one row: every approximate score is the same, nothing can be verified from the shadow;
    the retry ladder (fp32 rows, then the exact path) still returns the reference's insertion-order ties"""
    row = synth.make_corpus(1, 384, dup_frac=0.0)[0]
    corpus = np.tile(row, (2000, 1))
    g, o, ids = build_pair(corpus)
    res = g.search(row, 10)
    assert [r.node_id for r in res] == [ids[i].tobytes() for i in range(10)]
    assert_batch_equal(g, o, row[None, :], 10)
    # a cloud of rows within 1e-4 of each other around the query: closer together than the shadow's error bound
    rng = np.random.default_rng(5)
    base = synth.make_corpus(6000, 384, seed=9)
    q = base[17].copy()
    for r in range(100, 400):
        base[r] = q + rng.normal(0, 2e-4, 384).astype(np.float32)
    g, o, _ = build_pair(base)
    both_ways(g, o, q[None, :], 10, expect_half=False)


@pytest.mark.timeout(600)
def test_every_row_beats_the_cut_off():
    """Rows ordered by ASCENDING similarity to the query: every row a CTA meets clears its running cut-off, so the
    shared-memory lists fill at the highest possible rate between two checks (the worst case of the capacity
    argument of the one-barrier checkpoints, with the long check interval).  The best rows come last; a dropped
    append would lose exactly them."""
    rng = np.random.default_rng(11)
    n, d = 300_000, 384
    q = rng.standard_normal(d).astype(np.float32)
    q /= np.linalg.norm(q)
    noise = rng.standard_normal((n, d)).astype(np.float32)
    noise -= np.outer(noise @ q, q)                      # orthogonal to q
    noise /= np.linalg.norm(noise, axis=1, keepdims=True)
    # less and less noise: cosine 0.32 ... 0.69 over the bulk, then 150 well separated rows up to 0.999 (the
    # results must be verifiable from the shadow's approximate scores, or the call falls back to a tighter pass)
    w = np.concatenate([np.linspace(3.0, 1.05, n - 150), np.linspace(1.0, 0.05, 150)]).astype(np.float32)[:, None]
    corpus = (q[None, :] + w * noise).astype(np.float32)
    g, o, _ = build_pair(corpus)
    Q = np.stack([q, (q + 0.01 * noise[5]).astype(np.float32)])
    both_ways(g, o, Q[:1], 10)
    both_ways(g, o, Q, 10)
    both_ways(g, o, Q[:1], 100)
    # the same corpus through the tensor pass (private lists per epilogue thread, warp-cooperative compaction)
    Q8 = np.stack([(q + 0.01 * noise[i]).astype(np.float32) for i in range(8)])
    s0 = g.stats()
    assert_batch_equal(g, o, Q8, 10)
    assert g.stats()["queries_tensor"] - s0["queries_tensor"] == 8


def test_degenerate_queries():
    corpus = synth.make_corpus(5000, 384, seed=3)
    g, o, _ = build_pair(corpus)
    Q = synth.make_queries(corpus, 4, seed=3)
    Q[1] = 0.0
    Q[2, 5] = np.nan
    Q[3] *= np.float32(1e-30)
    assert_batch_equal(g, o, Q, 10)
    g.set_option("stream_bf16", 0)
    assert_batch_equal(g, o, Q, 10)


def test_index_without_shadow_streams_fp32():
    corpus = synth.make_corpus(4000, 384, seed=8)
    Q = synth.make_queries(corpus, 2, seed=8)
    ids = synth.make_ids(4000)
    g = GpuVectorIndex(384)
    g.set_option("shadow", 0)
    g.insert_batch(ids, corpus)
    o = OracleIndex(384, faithful_copy=False)
    o.insert_batch(ids, corpus)
    assert_batch_equal(g, o, Q, 10)
    st = g.stats()
    assert st["queries_stream_bf16"] == 0 and st["queries_stream"] == 2, st
