"""The restated HNSW branch (oracle/hnsw_oracle.c; parity unpinned, parameters unverified):
it must behave like an ANN index -- high recall against the exact oracle scan, and the
reference's own small-case assertions (vector/index.rs:484-535) hold through it too."""
import numpy as np

from cortex_b200 import synth
from oracle.binding import OracleHnsw, OracleIndex


def test_hnsw_small_cases_match_exact():
    v = np.array([[1.0, 0.0, 0.0], [0.9, 0.1, 0.0], [0.0, 1.0, 0.0]], np.float32)
    h = OracleHnsw(v)
    ids, d = h.search(np.array([1.0, 0.0, 0.0], np.float32), 2)
    assert len(ids) == 2 and ids[0] == 0      # test_index_insert_and_search
    assert d[0] == 0.0


def test_hnsw_recall_against_exact_scan():
    n, dim, b, k = 4000, 64, 50, 10
    corpus = synth.make_corpus(n, dim, seed=99)
    Q = synth.make_queries(corpus, b, seed=99)
    o = OracleIndex(dim, faithful_copy=False)
    o.insert_batch(synth.make_ids(n), corpus)
    h = OracleHnsw(corpus)
    hit = 0
    for q in Q:
        exact = o.search(q, k)
        ids, d = h.search(q, k)
        assert np.all(np.diff(d) >= 0)            # ascending distance
        # scores of whatever comes back are reference arithmetic: compare with the exact scan's values
        lut = {int(r): float(x) for r, x in zip(exact.rows, exact.distance)}
        for r, x in zip(ids, d):
            if int(r) in lut:
                assert np.float32(lut[int(r)]) == np.float32(x)
        hit += len(set(int(r) for r in ids) & set(int(r) for r in exact.rows))
    recall = hit / (b * k)
    assert recall >= 0.9, recall


def test_cpu_fast_courtesy_scan_agrees_with_the_oracle():
    """oracle/cpu_fast.c (the optimised, non-reference CPU baseline) returns the exact scan's ids on data
    without near-ties and scores within fp32 rounding of the reference's."""
    from oracle.binding import CpuFastScan

    n, dim, b, k = 6000, 96, 33, 10
    corpus = synth.make_corpus(n, dim, dup_frac=0.0, seed=17)
    Q = synth.make_queries(corpus, b, seed=18)
    o = OracleIndex(dim, faithful_copy=False)
    o.insert_batch(synth.make_ids(n), corpus)
    _, osc, _, orow, on = o.search_batch(Q, k)
    rows, score = CpuFastScan(corpus).search_batch(Q, k)
    same = sum(len(set(map(int, rows[i])) & set(map(int, orow[i]))) for i in range(b))
    assert same >= 0.99 * b * k
    assert np.max(np.abs(score - osc)) < 2e-6
