"""The scan step of AutoLinker::run_cycle (linker/auto_linker.rs:215-264) on the GPU against
the same loop restated over the CPU oracle: per new node search(emb, 100), skip self, keep
score >= auto_link_threshold (SimilarityLinkRule, linker/rules.rs:50), stop at
max_edges_per_node (auto_linker.rs:261)."""
import numpy as np
import pytest

from cortex_b200 import GpuVectorIndex, SimilarityConfig, synth
from oracle.binding import OracleIndex

pytestmark = pytest.mark.gpu


def reference_cycle(o, new_nodes, threshold, k, max_edges):
    out = {}
    for nid, emb in new_nodes:
        hits = o.search(emb, k)                      # auto_linker.rs:220-222
        edges = []
        for i in range(len(hits.score)):
            if hits.ids[i].tobytes() == nid:         # skip self, :235-237
                continue
            if hits.score[i] >= np.float32(threshold):   # rules.rs:50
                edges.append((hits.ids[i].tobytes(), float(hits.score[i])))
            if len(edges) >= max_edges:              # :261
                break
        out[nid] = edges
    return out


@pytest.mark.parametrize("n,b,max_edges", [(6000, 40, 50), (3000, 7, 3), (20_000, 300, 50)])
def test_autolink_batch_matches_reference_loop(n, b, max_edges):
    cfg = SimilarityConfig()
    corpus = synth.make_corpus(n, 384, n_clusters=max(4, n // 200), dup_frac=0.02, seed=n)
    ids = synth.make_ids(n)
    g = GpuVectorIndex(384)
    g.insert_batch(ids, corpus)
    o = OracleIndex(384, faithful_copy=False)
    o.insert_batch(ids, corpus)
    # new nodes: half already inserted (must skip themselves), half not yet in the index
    rng = np.random.default_rng(n)
    pick = rng.choice(n, size=b, replace=False)
    new_nodes = []
    for j, r in enumerate(pick):
        if j % 2 == 0:
            new_nodes.append((ids[r].tobytes(), corpus[r]))
        else:
            e = corpus[r] + rng.standard_normal(384).astype(np.float32) * 0.02
            new_nodes.append((synth.make_ids(1, start=10_000_000 + j)[0].tobytes(), (e / np.linalg.norm(e)).astype(np.float32)))
    got = g.autolink_batch(new_nodes, threshold=cfg.auto_link_threshold, k=100, max_edges_per_node=max_edges)
    exp = reference_cycle(o, new_nodes, cfg.auto_link_threshold, 100, max_edges)
    assert set(got) == set(exp)
    n_links = 0
    for nid in exp:
        assert [t for t, _ in got[nid]] == [t for t, _ in exp[nid]]
        assert np.array_equal(np.array([s for _, s in got[nid]], np.float32).view(np.uint32),
                              np.array([s for _, s in exp[nid]], np.float32).view(np.uint32))
        n_links += len(exp[nid])
    assert n_links > 0, "synthetic data produced no link candidates; the test would be vacuous"
