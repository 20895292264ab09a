"""Parity at the sizes BASELINE.json is quoted on.  The oracle restates the reference's algorithm
(per-pair scalar reductions), so it checks a SAMPLE of the queries of each batch against the full
corpus; the GPU path answers every query of the batch in one call.

  configs[1]  1M x 384, top-10: B=1 and B=4 (streaming pass), B=1024 (tensor pass)
  configs[3]  one shard's shape: 1024-d, top-100, B=256, 500k rows
  configs[4]  streaming ingest: 256-node batches searched (k=100, threshold 0.75) and appended to a
              growing corpus, across several extensions of the store
  plus size-independent properties at full size (self-match, sortedness, tie order)."""
import numpy as np
import pytest

from cortex_b200 import GpuVectorIndex, synth
from oracle.binding import OracleIndex

from _util import assert_batch_equal, same_bits

pytestmark = pytest.mark.gpu


def torch_corpus(n, d, seed):
    import torch

    import bench

    c = bench.make_corpus_torch(n, d, seed, torch.device("cuda", 0))
    q = bench.make_queries_torch(c, 1024, seed)
    return c, q


@pytest.fixture(scope="module")
def million():
    import torch

    n, d = 1_000_000, 384
    c, q = torch_corpus(n, d, 0xC027E5)
    ids = synth.make_ids(n)
    g = GpuVectorIndex(d)
    g.insert_batch_device(ids, c)
    corpus = c.cpu().numpy()
    Q = q.cpu().numpy()
    del c, q
    torch.cuda.empty_cache()
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(ids, corpus)
    return g, o, corpus, Q, ids


@pytest.mark.timeout(600)
def test_cfg2_1m_streaming_pass_b1_b4(million):
    g, o, corpus, Q, ids = million
    st0 = g.stats()
    assert_batch_equal(g, o, Q[:1], 10)
    assert_batch_equal(g, o, Q[1:3], 10)
    st1 = g.stats()
    assert st1["queries_stream"] - st0["queries_stream"] == 3, (st0, st1)
    assert st1["queries_stream_bf16"] - st0["queries_stream_bf16"] == 3, (st0, st1)  # over the bf16 shadow
    # the same from the fp32 rows, and four queries per pass (what the retry ladder and shadow-less indexes use)
    g.set_option("stream_bf16", 0)
    g.set_option("force_path", 1)
    try:
        assert_batch_equal(g, o, Q[:1], 10)
        assert_batch_equal(g, o, Q[1:5], 10)
    finally:
        g.set_option("stream_bf16", 1)
        g.set_option("force_path", 0)
    st2 = g.stats()
    assert st2["queries_stream"] - st1["queries_stream"] == 5 and st2["queries_stream_bf16"] == st1["queries_stream_bf16"]
    assert_batch_equal(g, o, Q[5:9], 10)   # B = 4: the tensor pass
    assert g.stats()["queries_tensor"] - st2["queries_tensor"] == 4


@pytest.mark.timeout(600)
def test_cfg2_1m_tensor_pass_b1024_sampled(million):
    g, o, corpus, Q, ids = million
    st0 = g.stats()
    sample = np.arange(0, 1024, 43)  # 24 of the 1024 queries go through the oracle
    assert_batch_equal(g, o, Q, 10, sample=sample)
    st1 = g.stats()
    assert st1["queries_tensor"] - st0["queries_tensor"] >= 1000, (st0, st1)
    # size-independent properties over ALL 1024 results
    gi, gs, gd, gn = g.search_batch_arrays(Q, 10)
    assert np.all(gn == 10)
    assert np.all(gs[:, :-1] >= gs[:, 1:]), "scores must be sorted descending"
    assert np.all((gs >= 0) & (gs <= 1))
    assert same_bits(np.clip(np.float32(1) - gd, 0, 1), gs), "score = clamp(1 - distance) (index.rs:254-256)"


@pytest.mark.timeout(600)
@pytest.mark.parametrize("b", [1024, 896, 640, 384])
def test_cfg2_1m_tensor_pass_leftover_sms(million, b):
    """Query-tile counts that do not divide the SM count (8, 7, 5, 3 tiles on 148 SMs): the CTAs left over by the
    (query tiles x row splits) grid walk the tail rows of several query tiles one after the other.  Same results
    as with those SMs idle, and as the oracle on a sample."""
    g, o, corpus, Q, ids = million
    q = Q[:b]
    a = g.search_batch_arrays(q, 10)
    g.set_option("tensor_leftover_sms", 0)
    try:
        ref = g.search_batch_arrays(q, 10)
    finally:
        g.set_option("tensor_leftover_sms", 1)
    assert np.array_equal(a[0], ref[0]) and same_bits(a[1], ref[1]) and same_bits(a[2], ref[2])
    assert np.array_equal(a[3], ref[3])
    # queries of the first and the last tile, and of tiles a left-over CTA switches between
    sample = np.unique(np.concatenate([np.arange(0, b, 97), [b - 1]]))
    assert_batch_equal(g, o, q, 10, sample=sample)


@pytest.mark.timeout(600)
def test_cfg2_1m_threshold_scan_with_leftover_sms(million):
    """search_threshold for 384 queries at 0.75 over 1 M rows: the tensor pass with a fixed cut-off, three query
    tiles (49 row splits + one left-over CTA that walks the tail rows for all three); 12 sampled queries against
    the reference loop (index.rs:376-388), the rest against the run with the left-over SMs idle."""
    g, o, corpus, Q, ids = million
    q = Q[:384]
    st0 = g.stats()
    a = g.search_threshold_batch_arrays(q, 0.75, 256)
    st1 = g.stats()
    assert st1["queries_tensor"] - st0["queries_tensor"] >= 380, (st0, st1)
    g.set_option("tensor_leftover_sms", 0)
    try:
        ref = g.search_threshold_batch_arrays(q, 0.75, 256)
    finally:
        g.set_option("tensor_leftover_sms", 1)
    for x, y in zip(a, ref):
        assert np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8))
    gi, gs, gd, gn, gt = a
    assert int(gt.max()) > 0, "the corpus is clustered: some queries must have partners above 0.75"
    for b in range(0, 384, 33):
        exp = o.search_threshold(q[b], 0.75)
        m = min(256, len(exp.ids))
        assert int(gt[b]) == len(exp.ids) and int(gn[b]) == m
        assert np.array_equal(gi[b, :m], exp.ids[:m]) and same_bits(gs[b, :m], exp.score[:m])
        assert same_bits(gd[b, :m], exp.distance[:m])


@pytest.mark.timeout(600)
def test_cfg2_1m_self_match_and_tie_order(million):
    """Rows of the corpus as queries: the row itself (cosine 1 up to rounding) and its exact duplicates
    lead the list, duplicates in insertion order."""
    g, o, corpus, Q, ids = million
    pick = np.arange(5, 1_000_000, 39_989)[:25]
    gi, gs, gd, gn = g.search_batch_arrays(corpus[pick], 10)
    for j, r in enumerate(pick):
        assert gs[j, 0] > 0.9999
        top = [int.from_bytes(gi[j, i, 8:].tobytes(), "big") for i in range(10)]
        assert r in top[:4]
    assert_batch_equal(g, o, corpus[pick[:6]], 10)


@pytest.mark.timeout(900)
def test_cfg4_shape_1024d_top100_b256():
    import torch

    n, d = 500_000, 1024
    c, q = torch_corpus(n, d, 77)
    # the corpus of configs[3] is bf16: rows are bf16 values (exactly representable in fp32)
    c = c.to(torch.bfloat16).to(torch.float32)
    ids = synth.make_ids(n)
    g = GpuVectorIndex(d)
    g.insert_batch_device(ids, c)
    corpus = c.cpu().numpy()
    Q = q[:256].cpu().numpy()
    del c, q
    torch.cuda.empty_cache()
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(ids, corpus)
    st0 = g.stats()
    assert_batch_equal(g, o, Q, 100, sample=np.arange(0, 256, 32))
    st1 = g.stats()
    assert st1["queries_tensor"] - st0["queries_tensor"] >= 240, (st0, st1)


@pytest.mark.timeout(900)
def test_cfg5_grow_while_searching():
    """256-node batches: auto-link scan against the corpus so far, then append.  The store is extended
    in place several times on the way; every batch is checked against the oracle, which grows alongside."""
    from test_gpu_autolink import reference_cycle

    d, n_batches = 384, 40
    corpus = synth.make_corpus(256 * n_batches, d, n_clusters=80, dup_frac=0.02, seed=55)
    ids = synth.make_ids(corpus.shape[0])
    g = GpuVectorIndex(d)
    o = OracleIndex(d, faithful_copy=False)
    grow0 = g.stats()["grow_events"]
    for bi in range(n_batches):
        lo, hi = 256 * bi, 256 * (bi + 1)
        if bi:
            nodes = [(ids[r].tobytes(), corpus[r]) for r in range(lo, hi)]
            got = g.autolink_batch(nodes, threshold=0.75, k=100, max_edges_per_node=50)
            if bi in (1, 2, 5, 17, 18, 39):  # the oracle loop is slow: check a few batches, incl. those around extensions
                exp = reference_cycle(o, nodes[::16], 0.75, 100, 50)
                for nid in exp:
                    assert [t for t, _ in got[nid]] == [t for t, _ in exp[nid]], f"batch {bi}"
                    assert same_bits([s for _, s in got[nid]], [s for _, s in exp[nid]])
        g.insert_batch(ids[lo:hi], corpus[lo:hi])
        o.insert_batch(ids[lo:hi], corpus[lo:hi])
    st = g.stats()
    assert st["grow_events"] - grow0 >= 2, st
    assert st["capacity_rows"] >= 256 * n_batches
    Q = synth.make_queries(corpus, 9, seed=3)
    assert_batch_equal(g, o, Q, 100)
    assert_batch_equal(g, o, Q[:3], 10)
