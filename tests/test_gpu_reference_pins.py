"""The reference's own unit tests for the vector index (vector/index.rs:475-729),
ported one to one and run against the CUDA path through the C ABI."""
import uuid

import numpy as np
import pytest

from cortex_b200 import CortexError, GpuVectorIndex, SimilarityConfig, VectorFilter

pytestmark = pytest.mark.gpu


def nid() -> bytes:
    return uuid.uuid4().bytes


def test_index_insert_and_search():  # index.rs:484-510
    ix = GpuVectorIndex(3)
    id1, id2, id3 = nid(), nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.9, 0.1, 0.0])
    ix.insert(id3, [0.0, 1.0, 0.0])
    ix.rebuild()
    r = ix.search([1.0, 0.0, 0.0], 2)
    assert len(r) == 2 and r[0].node_id == id1


def test_threshold_search():  # :513-535
    ix = GpuVectorIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.0, 1.0, 0.0])
    ix.rebuild()
    r = ix.search_threshold([1.0, 0.0, 0.0], 0.95)
    assert len(r) == 1 and r[0].node_id == id1


def test_index_persistence(tmp_path):  # :538-566
    p = str(tmp_path / "test.hnsw")
    ix = GpuVectorIndex(3)
    id1 = nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.rebuild()
    ix.save(p)
    ld = GpuVectorIndex.load(p)
    assert len(ld) == 1
    r = ld.search([1.0, 0.0, 0.0], 1)
    assert len(r) == 1 and r[0].node_id == id1


def test_dimension_mismatch_rejected():  # :579-583
    ix = GpuVectorIndex(3)
    with pytest.raises(CortexError) as e:
        ix.insert(nid(), [1.0, 2.0])
    assert "Embedding dimension mismatch: expected 3, got 2" in str(e.value)


def test_empty_index_search():  # :586-590
    ix = GpuVectorIndex(3)
    assert ix.search([1.0, 0.0, 0.0], 5) == []
    assert ix.is_empty()


def test_brute_force_fallback():  # :593-606
    ix = GpuVectorIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.0, 1.0, 0.0])
    r = ix.search([1.0, 0.0, 0.0], 2)
    assert len(r) == 2 and r[0].node_id == id1


def test_filter_by_kind():  # :609-627
    ix = GpuVectorIndex.with_metadata(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.set_metadata(id1, "fact", "test")
    ix.insert(id2, [0.9, 0.1, 0.0])
    ix.set_metadata(id2, "decision", "test")
    ix.rebuild()
    r = ix.search([1.0, 0.0, 0.0], 5, VectorFilter.new().with_kinds(["decision"]))
    assert len(r) == 1 and r[0].node_id == id2


def test_filter_exclude():  # :630-646
    ix = GpuVectorIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.9, 0.1, 0.0])
    ix.rebuild()
    r = ix.search([1.0, 0.0, 0.0], 5, VectorFilter.new().excluding([id1]))
    assert len(r) == 1 and r[0].node_id == id2


def test_remove_doesnt_crash_search():  # :649-664
    ix = GpuVectorIndex(3)
    id1, id2 = nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.0, 1.0, 0.0])
    ix.rebuild()
    ix.remove(id1)
    assert len(ix) == 1
    assert len(ix.search([1.0, 0.0, 0.0], 5)) > 0
    ix.remove(nid())  # unknown id is not an error (index.rs:316-323)


def test_search_batch():  # :667-684
    ix = GpuVectorIndex(3)
    id1, id2, id3 = nid(), nid(), nid()
    ix.insert(id1, [1.0, 0.0, 0.0])
    ix.insert(id2, [0.0, 1.0, 0.0])
    ix.insert(id3, [0.0, 0.0, 1.0])
    ix.rebuild()
    res = ix.search_batch([(id1, [1.0, 0.0, 0.0]), (id2, [0.0, 1.0, 0.0])], 1)
    assert len(res) == 2
    assert res[id1][0].node_id == id1 and res[id2][0].node_id == id2


def test_similarity_score_range():  # :687-708
    ix = GpuVectorIndex(3)
    ix.insert(nid(), [1.0, 0.0, 0.0])
    ix.insert(nid(), [-1.0, 0.0, 0.0])
    ix.rebuild()
    r = ix.search([1.0, 0.0, 0.0], 2)
    assert all(0.0 <= x.score <= 1.0 for x in r)
    assert r[0].score > 0.99
    assert r[1].distance == 2.0 and r[1].score == 0.0


def test_threshold_returns_only_above():  # :711-728
    ix = GpuVectorIndex(3)
    idc, idf = nid(), nid()
    ix.insert(idc, [1.0, 0.0, 0.0])
    ix.insert(idf, [0.0, 0.0, 1.0])
    ix.rebuild()
    r = ix.search_threshold([1.0, 0.0, 0.0], 0.5)
    assert all(x.score >= 0.5 for x in r)
    assert any(x.node_id == idc for x in r)


def test_zero_norm_row_is_nan_and_last():
    ix = GpuVectorIndex(3)
    a, z = nid(), nid()
    ix.insert(z, [0.0, 0.0, 0.0])
    ix.insert(a, [0.0, 1.0, 0.0])
    r = ix.search([1.0, 0.0, 0.0], 2)
    assert r[0].node_id == a and r[0].score == 0.0
    assert r[1].node_id == z and np.isnan(r[1].score) and np.isnan(r[1].distance)
    assert [x.node_id for x in ix.search_threshold([1.0, 0.0, 0.0], 0.0)] == [a]
