"""The embedding store and the call shapes around it: in-place growth, upserts inside one batch,
asynchronous mutations, dedup fallbacks, graph replay of repeated device-resident searches, the
irregular-row count."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cortex_b200 import CortexError, GpuVectorIndex, synth
from oracle.binding import OracleIndex

from _util import assert_batch_equal, assert_dedup_equal, build_pair, check_threshold_batch, same_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_store_grows_in_place_and_keeps_results():
    corpus = synth.make_corpus(30_000, 384, seed=2)
    ids = synth.make_ids(30_000)
    g = GpuVectorIndex(384)
    o = OracleIndex(384, faithful_copy=False)
    Q = synth.make_queries(corpus, 6, seed=1)
    at = 0
    for step in (3, 250, 1, 4000, 700, 25_046):
        g.insert_batch(ids[at:at + step], corpus[at:at + step])
        o.insert_batch(ids[at:at + step], corpus[at:at + step])
        at += step
        assert_batch_equal(g, o, Q, 10)
    st = g.stats()
    assert st["in_place_growth"] == 1, "the driver's virtual-memory API should be available on a B200 box"
    assert st["grow_events"] >= 3 and st["capacity_rows"] >= 30_000


def test_growth_fallback_without_virtual_memory_api():
    """CORTEX_GPU_NO_VMM=1 forces the allocate-copy-free path; same results."""
    code = """
import numpy as np
from cortex_b200 import GpuVectorIndex, synth
c = synth.make_corpus(9000, 128, seed=4); ids = synth.make_ids(9000)
g = GpuVectorIndex(128)
for a in range(0, 9000, 1500):
    g.insert_batch(ids[a:a+1500], c[a:a+1500])
st = g.stats(); assert st['in_place_growth'] == 0 and st['grow_events'] >= 2, st
r = g.search(c[1234], 3); assert r[0].node_id == ids[1234].tobytes() or r[0].score > 0.9999
g.remove(ids[7].tobytes()); g.rebuild(); assert len(g) == 8999
print('ok')
"""
    env = dict(os.environ, CORTEX_GPU_NO_VMM="1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=170)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_upserts_and_duplicates_inside_one_batch():
    """HashMap::insert (index.rs:307): an id that appears twice keeps its first slot and its last
    vector; ids that already exist are overwritten in place."""
    rng = np.random.default_rng(0)
    d = 64
    base = rng.standard_normal((600, d)).astype(np.float32)
    ids = synth.make_ids(600)
    g = GpuVectorIndex(d)
    o = OracleIndex(d, faithful_copy=False)
    g.insert_batch(ids[:300], base[:300])
    o.insert_batch(ids[:300], base[:300])
    # one batch: new rows, an existing id, a new id twice, an existing id twice
    order = [300, 301, 5, 302, 302, 303, 7, 7, 304]
    rows = rng.standard_normal((len(order), d)).astype(np.float32)
    g.insert_batch(ids[order], rows)
    for j, r in enumerate(order):
        o.insert(ids[r].tobytes(), rows[j])
    assert len(g) == len(o) == 305
    Q = np.concatenate([rows, base[:4]])
    assert_batch_equal(g, o, Q, 20)
    # a failing insert leaves the index untouched
    with pytest.raises(CortexError):
        g.insert_batch(ids[400:402], rng.standard_normal((2, d + 1)).astype(np.float32))
    assert len(g) == 305
    assert_batch_equal(g, o, Q[:3], 20)


def test_device_batch_insert_is_an_upsert():
    import torch

    d = 128
    corpus = synth.make_corpus(5000, d, seed=6)
    ids = synth.make_ids(5000)
    g = GpuVectorIndex(d)
    o = OracleIndex(d, faithful_copy=False)
    g.insert_batch_device(ids[:4000], torch.from_numpy(corpus[:4000]).cuda())
    o.insert_batch(ids[:4000], corpus[:4000])
    # second batch overlaps the first: rows 3990..3999 are overwritten with new vectors, 4000.. appended
    newer = corpus[3990:5000].copy()
    newer[:10] = corpus[10:20]
    g.insert_batch_device(ids[3990:5000], torch.from_numpy(newer).cuda())
    for j, r in enumerate(range(3990, 5000)):
        o.insert(ids[r].tobytes(), newer[j])
    assert len(g) == len(o) == 5000
    assert_batch_equal(g, o, synth.make_queries(corpus, 8, seed=2), 10)
    assert_batch_equal(g, o, corpus[10:14], 5)


def test_mutations_return_before_the_device_finishes_and_searches_see_them():
    corpus = synth.make_corpus(20_000, 384, seed=9)
    g, o, ids = build_pair(corpus)
    Q = synth.make_queries(corpus, 5, seed=3)
    for r in range(0, 2000, 3):
        g.remove(ids[r].tobytes())
        o.remove(ids[r].tobytes())
    for r in range(1, 2000, 5):
        v = corpus[(r * 13) % 20_000] * np.float32(0.5)
        g.insert(ids[r].tobytes(), v)
        o.insert(ids[r].tobytes(), v)
    assert_batch_equal(g, o, Q, 10)
    g.rebuild()
    assert_batch_equal(g, o, Q, 10)
    assert len(g) == len(o)


def test_dedup_scan_small_index_and_low_threshold_fall_back_to_exact():
    """The reference scan always succeeds (dedup.rs:65-127): fewer than 256 rows and thresholds below the
    fast path's floor are served by the exact path."""
    corpus = synth.make_corpus(120, 32, n_clusters=6, seed=3)
    g, o, ids = build_pair(corpus)
    exp = assert_dedup_equal(g, o, corpus, ids, 0.9, per_node_cap=120)
    assert len(exp) > 0
    assert_dedup_equal(g, o, corpus, ids, 0.05, per_node_cap=120)
    corpus2 = synth.make_corpus(700, 48, n_clusters=9, seed=4)
    g2, o2, ids2 = build_pair(corpus2)
    assert_dedup_equal(g2, o2, corpus2, ids2, 0.1, per_node_cap=700)


def test_dedup_scan_large_duplicate_cluster():
    """3000 copies of one embedding -- the typical dedup case -- overflow the fast path's per-node lists
    (2048); those nodes take the exact path and the scan still returns every pair."""
    d = 32
    corpus = synth.make_corpus(4000, d, seed=8)
    corpus[500:3500] = corpus[3]
    g, o, ids = build_pair(corpus)
    a, b, sc, total = g.dedup_scan(0.99, per_node_cap=8, max_pairs=200_000)
    # every pair inside the cluster {3} U [500, 3500) qualifies: m (m - 1) / 2 with m = 3001
    m = 3001
    assert total >= m * (m - 1) // 2
    # per node at most 8 partners are written, the best first; for the first cluster member they are the next 8 copies
    first = [(a[i].tobytes(), b[i].tobytes()) for i in range(len(a)) if a[i].tobytes() == ids[3].tobytes()]
    assert [p[1] for p in first] == [ids[r].tobytes() for r in range(500, 508)]
    assert g.stats()["queries_exact"] > 0


def test_dedup_scan_with_irregular_row_still_succeeds():
    corpus = synth.make_corpus(600, 16, n_clusters=5, seed=5)
    corpus[100] = np.float32(1e-25)  # norm underflows fp32: outside the fast passes' error bounds
    g, o, ids = build_pair(corpus)
    assert g.stats()["irregular_rows"] == 1
    assert_dedup_equal(g, o, corpus, ids, 0.9, per_node_cap=600)
    # removing the row restores the fast passes without a rebuild
    g.remove(ids[100].tobytes())
    o.remove(ids[100].tobytes())
    assert g.stats()["irregular_rows"] == 0
    st0 = g.stats()
    assert_batch_equal(g, o, corpus[:2], 5)
    assert g.stats()["queries_stream"] > st0["queries_stream"]


def test_repeated_device_searches_replay_a_graph():
    import torch

    corpus = synth.make_corpus(60_000, 384, seed=31)
    g, o, ids = build_pair(corpus)
    Q = synth.make_queries(corpus, 512, seed=7)
    dq = torch.from_numpy(Q).cuda()
    s = torch.cuda.current_stream().cuda_stream
    oi, os_, od, orow, on = o.search_batch(Q[::32], 10)
    out = None
    for profile in (0, 1):
        g.set_option("profile", profile)
        st0 = g.stats()
        for it in range(5):
            out = g.search_batch_device(dq, 10, stream=s, out=out)
            torch.cuda.synchronize()
            rows, sc, di, n = (t.cpu().numpy() for t in out)
            assert np.array_equal(rows[::32], orow) and same_bits(sc[::32], os_) and same_bits(di[::32], od)
        st1 = g.stats()
        assert st1["graph_launches"] - st0["graph_launches"] >= 3, (profile, st0, st1)
        if profile:
            assert st1["pass_kernel_ns"] > st0["pass_kernel_ns"]
    # a mutation changes the shape: the recording is dropped, results follow the new corpus
    g.remove(ids[int(orow[0, 0])].tobytes())
    o.remove(ids[int(orow[0, 0])].tobytes())
    out = g.search_batch_device(dq, 10, stream=s, out=out)
    torch.cuda.synchronize()
    _, _, _, orow2, _ = o.search_batch(Q[:1], 10)
    assert np.array_equal(out[0][:1].cpu().numpy(), orow2)
    g.set_option("graphs", 0)
    st0 = g.stats()
    for _ in range(3):
        out = g.search_batch_device(dq, 10, stream=s, out=out)
    assert g.stats()["graph_launches"] == st0["graph_launches"]


def test_near_ties_below_one_half():
    """Scores below 0.5 are quantised to 2^-24 by 1 - (1 - s) (index.rs:177,255): many rows whose cosines
    differ by less than that tie, and ties straddle rank k.  Both fast passes must hand such queries to a
    tighter path rather than emit an order the reference would not."""
    rng = np.random.default_rng(12)
    d, n = 384, 6000
    base = rng.standard_normal(d).astype(np.float32)
    base /= np.linalg.norm(base)
    ortho = rng.standard_normal((n, d)).astype(np.float32)
    ortho -= np.outer(ortho @ base, base)
    ortho /= np.linalg.norm(ortho, axis=1, keepdims=True)
    # cosine with `base` ~ 0.3 +- 3e-8 for a block of rows: distinct cosines, (mostly) equal scores
    eps = rng.uniform(-3e-8, 3e-8, n).astype(np.float64)
    c = 0.3 + np.where(np.arange(n) % 3 == 0, eps, rng.uniform(-0.2, 0.2, n))
    corpus = (np.outer(c, base) + np.sqrt(1 - c * c)[:, None] * ortho).astype(np.float32)
    g, o, ids = build_pair(corpus)
    Q = np.stack([base, base * 2.5, (base + 0.001 * ortho[0]).astype(np.float32)])
    for k in (1, 10, 37, 100):
        assert_batch_equal(g, o, Q, k)
    g.set_option("force_path", 2)   # the tensor pass nominates
    Q40 = np.concatenate([Q] * 14)[:40]
    for k in (10, 100):
        assert_batch_equal(g, o, Q40, k)
