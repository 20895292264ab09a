"""The oracle against every assertion the reference's own unit tests make for
the scan path (vector/index.rs:475-729), ported one to one, plus bit-for-bit
agreement with the independently written numpy golden vectors.

The same scenarios are run against the CUDA path in tests/test_gpu_reference_pins.py.
"""
import os
import uuid

import numpy as np
import pytest

from oracle.binding import Filter, OracleIndex, distance, distance_to_similarity

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "scan_golden.npz"))


def nid() -> bytes:
    return uuid.uuid4().bytes


# ---- vector/index.rs:484-510 test_index_insert_and_search -------------------
def test_index_insert_and_search():
    ix = OracleIndex(3)
    id1, id2, id3 = nid(), nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.9, 0.1, 0.0])
    ix.insert(id3, [0.0, 1.0, 0.0])
    ix.rebuild()
    r = ix.search([1.0, 0.0, 0.0], 2)
    assert len(r.score) == 2
    assert bytes(r.ids[0]) == id1


# ---- :513-535 test_threshold_search -----------------------------------------
def test_threshold_search():
    ix = OracleIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.0, 1.0, 0.0])
    ix.rebuild()
    r = ix.search_threshold([1.0, 0.0, 0.0], 0.95)
    assert len(r.score) == 1
    assert bytes(r.ids[0]) == id1


# ---- :538-566 test_index_persistence ----------------------------------------
def test_index_persistence(tmp_path):
    p = str(tmp_path / "test.hnsw")
    ix = OracleIndex(3)
    id1 = nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.rebuild()
    ix.save(p)
    ld = OracleIndex.load(p)
    assert len(ld) == 1
    r = ld.search([1.0, 0.0, 0.0], 1)
    assert len(r.score) == 1 and bytes(r.ids[0]) == id1


def test_save_layout_is_bincode(tmp_path):
    """bincode 1.3 fixint layout of (vectors, metadata, dimension), index.rs:437-445."""
    p = str(tmp_path / "ix.bin")
    ix = OracleIndex(3)
    i1 = bytes(range(16))
    ix.insert(i1, [1.0, 2.0, 3.0])
    ix.set_metadata(i1, "fact", "agent-a")
    ix.save(p)
    raw = open(p, "rb").read()
    exp = (np.uint64(1).tobytes() + np.uint64(16).tobytes() + i1 + np.uint64(3).tobytes()
           + np.array([1, 2, 3], np.float32).tobytes()
           + np.uint64(1).tobytes() + np.uint64(16).tobytes() + i1
           + np.uint64(4).tobytes() + b"fact" + np.uint64(7).tobytes() + b"agent-a"
           + np.uint64(3).tobytes())
    assert raw == exp


# ---- :579-583 test_dimension_mismatch_rejected -------------------------------
def test_dimension_mismatch_rejected():
    ix = OracleIndex(3)
    with pytest.raises(ValueError):
        ix.insert(nid(), [1.0, 2.0])


# ---- :586-590 test_empty_index_search ----------------------------------------
def test_empty_index_search():
    ix = OracleIndex(3)
    assert len(ix.search([1.0, 0.0, 0.0], 5).score) == 0


# ---- :593-606 test_brute_force_fallback --------------------------------------
def test_brute_force_fallback():
    ix = OracleIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.0, 1.0, 0.0])
    r = ix.search([1.0, 0.0, 0.0], 2)
    assert len(r.score) == 2 and bytes(r.ids[0]) == id1


# ---- :609-627 test_filter_by_kind --------------------------------------------
def test_filter_by_kind():
    ix = OracleIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.set_metadata(id1, "fact", "test")
    ix.insert(id2, [0.9, 0.1, 0.0])
    ix.set_metadata(id2, "decision", "test")
    ix.rebuild()
    r = ix.search([1.0, 0.0, 0.0], 5, Filter(kinds=["decision"]))
    assert len(r.score) == 1 and bytes(r.ids[0]) == id2


def test_filter_without_metadata_passes():
    """index.rs:233-248: ids with no metadata pass every kind / agent filter."""
    ix = OracleIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.9, 0.1, 0.0])
    ix.set_metadata(id2, "fact", "a1")
    r = ix.search([1.0, 0.0, 0.0], 5, Filter(kinds=["decision"], source_agent="zz"))
    assert [bytes(i) for i in r.ids] == [id1]


# ---- :630-646 test_filter_exclude --------------------------------------------
def test_filter_exclude():
    ix = OracleIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.9, 0.1, 0.0])
    ix.rebuild()
    r = ix.search([1.0, 0.0, 0.0], 5, Filter(exclude=[id1]))
    assert len(r.score) == 1 and bytes(r.ids[0]) == id2


# ---- :649-664 test_remove_doesnt_crash_search --------------------------------
def test_remove_doesnt_crash_search():
    ix = OracleIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.0, 1.0, 0.0])
    ix.rebuild()
    ix.remove(id1)
    assert len(ix) == 1
    assert len(ix.search([1.0, 0.0, 0.0], 5).score) > 0


# ---- :667-684 test_search_batch ----------------------------------------------
def test_search_batch():
    ix = OracleIndex(3)
    id1, id2, id3 = nid(), nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.0, 1.0, 0.0])
    ix.insert(id3, [0.0, 0.0, 1.0])
    ix.rebuild()
    ids, sc, di, rows, n = ix.search_batch(np.array([[1, 0, 0], [0, 1, 0]], np.float32), 1)
    assert list(n) == [1, 1]
    assert bytes(ids[0, 0]) == id1 and bytes(ids[1, 0]) == id2


# ---- :687-708 test_similarity_score_range ------------------------------------
def test_similarity_score_range():
    ix = OracleIndex(3)
    ix.insert(nid(), [1.0, 0.0, 0.0])
    ix.insert(nid(), [-1.0, 0.0, 0.0])
    ix.rebuild()
    r = ix.search([1.0, 0.0, 0.0], 2)
    assert all(0.0 <= s <= 1.0 for s in r.score)
    assert r.score[0] > 0.99
    assert r.distance[1] == 2.0 and r.score[1] == 0.0  # distance stays unclamped (index.rs:280-284)


# ---- :711-728 test_threshold_returns_only_above ------------------------------
def test_threshold_returns_only_above():
    ix = OracleIndex(3)
    idc, idf = nid(), nid()
    ix.insert(idc, [1.0, 0.0, 0.0])
    ix.insert(idf, [0.0, 0.0, 1.0])
    ix.rebuild()
    r = ix.search_threshold([1.0, 0.0, 0.0], 0.5)
    assert all(s >= 0.5 for s in r.score)
    assert any(bytes(i) == idc for i in r.ids)


def test_insert_same_id_overwrites():
    """HashMap::insert semantics, index.rs:307."""
    ix = OracleIndex(3)
    i = nid()
    ix.insert(i, [1.0, 0.0, 0.0])
    ix.insert(i, [0.0, 1.0, 0.0])
    assert len(ix) == 1
    r = ix.search([0.0, 1.0, 0.0], 1)
    assert r.score[0] == 1.0


def test_scalar_arithmetic_known_answers():
    assert distance([1, 0, 0], [1, 0, 0]) == 0.0
    assert distance([1, 0, 0], [0, 1, 0]) == 1.0
    assert distance([1, 0, 0], [-1, 0, 0]) == 2.0
    assert np.isnan(distance([1, 0, 0], [0, 0, 0]))  # 0/0, index.rs:176
    assert distance_to_similarity(2.0) == 0.0
    assert distance_to_similarity(-0.5) == 1.0
    assert np.isnan(distance_to_similarity(float("nan")))  # f32::clamp keeps NaN
    # zip truncation (index.rs:172): dot over 2 terms, norms over full slices
    d = distance([3.0, 4.0], [3.0, 4.0, 12.0])
    assert d == np.float32(1.0) - np.float32(25.0) / (np.float32(5.0) * np.float32(13.0))


@pytest.mark.parametrize("case", ["A", "B", "C", "D"])
def test_oracle_matches_numpy_golden_bitwise(case):
    corpus, query = GOLD[f"{case}_corpus"], GOLD[f"{case}_query"]
    ix = OracleIndex(corpus.shape[1])
    for r in range(corpus.shape[0]):
        ix.insert(r.to_bytes(16, "big"), corpus[r])
    qs = query if query.ndim == 2 else query[None, :]
    rows = GOLD[f"{case}_rows"].reshape(len(qs), -1)
    score = GOLD[f"{case}_score"].reshape(len(qs), -1)
    dist = GOLD[f"{case}_dist"].reshape(len(qs), -1)
    for b, q in enumerate(qs):
        h = ix.search(q, corpus.shape[0])
        assert np.array_equal(h.rows, rows[b])
        assert np.array_equal(h.score.view(np.uint32), score[b].view(np.uint32))
        assert np.array_equal(h.distance.view(np.uint32), dist[b].view(np.uint32))
        # truncation = prefix
        h5 = ix.search(q, 5)
        assert np.array_equal(h5.rows, rows[b][:5])
        # threshold = prefix of the sorted list with score >= t (NaN never passes)
        t = np.float32(0.75)
        ht = ix.search_threshold(q, float(t))
        exp = [r for r, s in zip(rows[b], score[b]) if s >= t]
        assert list(ht.rows) == exp
