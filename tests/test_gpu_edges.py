"""GPU parity on the edges of the call space: k at and beyond the fast path's limit, tiny and
tile-boundary corpora, empty and degenerate queries, extreme norms, ties that span the k-th
place inside a large batch.  Everything is compared with the CPU oracle bit for bit."""
import numpy as np
import pytest

from cortex_b200 import GpuVectorIndex, synth
from oracle.binding import OracleIndex

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


def check(g, o, Q, k):
    gi, gs, gd, gn = g.search_batch_arrays(Q, k)
    oi, os_, od, _, on = o.search_batch(Q, k)
    assert np.array_equal(gn, on), (gn, on)
    for b in range(Q.shape[0]):
        n = int(on[b])
        assert np.array_equal(gi[b, :n], oi[b, :n]), f"query {b}: ids differ"
        assert same_bits(gs[b, :n], os_[b, :n]) and same_bits(gd[b, :n], od[b, :n]), f"query {b}"


@pytest.mark.parametrize("k", [1, 16, 17, 48, 49, 100, 112, 113, 128, 129, 200, 5000])
@pytest.mark.parametrize("b", [2, 40])
def test_k_across_the_fast_path_limits(k, b):
    """keep counts change at k = 16/48/112 (tensor pass: 32/64/128 keys), the fast passes stop at
    k = 128, k > rows returns every row."""
    corpus = synth.make_corpus(3000, 128, seed=100 + k)
    Q = synth.make_queries(corpus, b, seed=k)
    g, o, _ = build_pair(corpus)
    check(g, o, Q, k)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 256, 257, 511, 512, 513])
def test_corpus_sizes_around_tile_boundaries(n):
    corpus = synth.make_corpus(n, 96, seed=n)
    Q = synth.make_queries(corpus, 7, seed=n + 1)
    g, o, _ = build_pair(corpus)
    check(g, o, Q, 10)
    check(g, o, Q[:1], 3)


def test_empty_batch_and_k_zero():
    corpus = synth.make_corpus(600, 64, seed=1)
    g, o, _ = build_pair(corpus)
    ids, sc, di, n = g.search_batch_arrays(np.zeros((0, 64), np.float32), 5)
    assert n.shape == (0,)
    assert g.search(corpus[0], 0) == []


def test_degenerate_queries():
    """zero query (every score NaN: rows come back in row order), NaN / inf components, huge and
    denormal magnitudes -- the oracle follows the reference's arithmetic and so must the GPU."""
    corpus = synth.make_corpus(4000, 64, zero_row=True, seed=2)
    g, o, _ = build_pair(corpus)
    z = np.zeros(64, np.float32)
    q_nan = corpus[3].copy(); q_nan[5] = np.nan
    q_inf = corpus[4].copy(); q_inf[7] = np.inf
    q_big = corpus[5] * np.float32(1e30)
    q_tiny = corpus[6] * np.float32(1e-30)
    for q in (z, q_nan, q_inf, q_big, q_tiny):
        check(g, o, q[None, :], 10)
    check(g, o, np.stack([z, q_nan, corpus[9], q_inf, q_big, q_tiny, corpus[11]]), 10)  # tensor pass + fallbacks


def test_extreme_row_norms():
    corpus = synth.make_corpus(3000, 64, seed=3)
    corpus[10] *= np.float32(1e25)     # norm overflows to inf: sim = x / inf
    corpus[11] *= np.float32(1e-25)    # squares underflow: norm 0 -> NaN score
    corpus[12] *= np.float32(1e18)
    corpus[13] *= np.float32(1e-18)
    Q = np.stack([corpus[10], corpus[11], corpus[12], corpus[13], corpus[14]])
    g, o, ids = build_pair(corpus)
    check(g, o, Q, 10)
    check(g, o, Q[:2], 10)
    # rows whose norm is outside [1e-15, 1e15] (squares under- / overflow fp32, or come close) are outside
    # the fast passes' error bounds: while the index holds them every search is served by the exact path ...
    st = g.stats()
    assert st["queries_exact"] == 7 and st["queries_tensor"] == 0 and st["queries_stream"] == 0, st
    # ... and once they are gone (removed + compacted) the fast passes are back
    for r in (10, 11, 12, 13):
        g.remove(ids[r].tobytes())
        o.remove(ids[r].tobytes())
    g.rebuild()
    check(g, o, Q, 10)   # four of these queries have irregular norms themselves and still go to the exact path
    R = synth.make_queries(corpus, 6, seed=8)
    check(g, o, R, 10)
    st = g.stats()
    assert st["queries_tensor"] + st["queries_stream"] >= 7, st


def test_ties_across_the_kth_place_in_a_large_batch():
    """300 identical rows straddle rank k: the reference's stable sort keeps insertion order, and so do
    the tensor pass + exact rescoring (or the fallback they hand over to)."""
    corpus = synth.make_corpus(6000, 128, seed=4)
    corpus[1000:1300] = corpus[17]
    Q = np.concatenate([np.tile(corpus[17], (3, 1)), synth.make_queries(corpus, 61, seed=5)])
    g, o, _ = build_pair(corpus)
    check(g, o, Q, 10)
    check(g, o, Q, 100)
