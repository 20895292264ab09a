"""GPU: the node extractor / bulk load and the score decay, through the C ABI, against oracle/aux_oracle.c."""
import os

import numpy as np
import pytest

from cortex_b200 import GpuVectorIndex, synth
from cortex_b200.index import extract_embeddings
from oracle.binding import OracleIndex, apply_score_decay, walk_node

from _nodes import canonical_node, node_bytes, pack
from _util import assert_batch_equal, same_bits

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "node_golden.bin")


def make_table(n, dim, seed=0):
    rng = np.random.default_rng(seed)
    corpus = synth.make_corpus(n, dim, seed=seed + 1)
    ids = synth.make_ids(n)
    values = []
    for i in range(n):
        kw = {}
        r = i % 17
        if r == 3:
            kw["embedding"] = None
        elif r == 5:
            kw["embedding"] = corpus[i][: dim - 1]
        else:
            kw["embedding"] = corpus[i]
        if r == 7:
            kw["deleted"] = True
        if r == 9:
            kw["metadata_raw"] = (1).to_bytes(8, "little") + (1).to_bytes(8, "little") + b"k" + (2).to_bytes(8, "little") + b"vv"
        secs = 1_700_000_000 + int(rng.integers(0, 5000))  # many equal timestamps: ties keep table order
        frac = ["", ".5", ".123", ".000001", ".999999999"][i % 5]
        t = np.datetime64(secs, "s").astype("datetime64[s]").astype(str) + frac + "Z"
        kw["created"] = t
        kw["title"] = "t" * (i % 40) + ("é" if i % 3 == 0 else "")
        kw["body"] = "body " * (i % 11)
        kw["tags"] = tuple(f"tag{j}" for j in range(i % 4))
        kw["session"] = "s1" if i % 2 else None
        kw["access_count"] = i
        values.append(node_bytes(ids[i].tobytes(), **kw))
    values[11] = values[11][:-25]            # truncated record
    values[12] = canonical_node()            # the reference's own fixture
    return values, corpus, ids


def test_extractor_matches_oracle_walk_and_golden_bytes():
    dim = 96
    values, corpus, ids = make_table(400, dim)
    blob, offs = pack(values)
    out = extract_embeddings(blob, offs, dim)
    seen = set()
    for i, v in enumerate(values):
        st, nid, row, created, last, acc = walk_node(v, dim)
        assert out["status"][i] == st, i
        seen.add(st)
        if st in (0, 1, 2, 3):
            assert out["ids"][i].tobytes() == nid and out["created_ns"][i] == created
            assert out["last_accessed_ns"][i] == last and out["access_count"][i] == acc
        if st == 0:
            assert same_bits(out["rows"][i], row), i
    assert seen == {0, 1, 2, 3, 4, 5}
    g = open(GOLD, "rb").read()
    o1 = extract_embeddings(np.frombuffer(g, np.uint8), np.array([0, len(g)], np.uint64), 384)
    assert o1["status"][0] == 1 and o1["created_ns"][0] == 1_700_000_000 * 10**9
    assert o1["ids"][0].tobytes() == bytes([0x01, 0x92, 0xab, 0xcd, 0xef, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11])


@pytest.mark.parametrize("devices", [None, [0, 0, 0]])
def test_load_nodes_is_the_startup_loop(devices):
    """serve.rs:105-123: list_nodes (newest first, stable) then insert every embedding; searches over the result
    equal the oracle index built by that loop."""
    dim = 128
    values, corpus, ids = make_table(3000, dim, seed=4)
    blob, offs = pack(values)
    g = GpuVectorIndex(dim, devices=devices) if devices else GpuVectorIndex(dim)
    status, counts = g.load_nodes(blob, offs)
    walked = [walk_node(v, dim) for v in values]
    assert [w[0] for w in walked] == status.tolist()
    assert counts.tolist() == [sum(1 for w in walked if w[0] == s) for s in range(6)]
    order = sorted([i for i, w in enumerate(walked) if w[0] == 0], key=lambda i: -walked[i][3])  # stable, created desc
    o = OracleIndex(dim, faithful_copy=False)
    for i in order:
        o.insert(walked[i][1], walked[i][2])
    assert len(g) == len(o) == len(order)
    Q = synth.make_queries(corpus, 12, seed=9)
    assert_batch_equal(g, o, Q, 20)
    assert_batch_equal(g, o, corpus[order[:3]], 5)   # exact duplicates of stored rows: tie order = insertion order


def test_score_decay_matches_oracle_and_reranks():
    rng = np.random.default_rng(3)
    n, seg = 6000, 300
    raw = rng.uniform(0, 1, n).astype(np.float32)
    idle = rng.integers(-100, 500 * 86400, n)
    acc = rng.integers(0, 200, n).astype(np.uint64)
    rate = rng.choice([0.05, 0.04, 0.005, 0.01, 0.02], n)
    raw[::50] = raw[1::50]   # equal raw scores
    idle[::50] = idle[1::50]
    acc[::50] = acc[1::50]
    rate[::50] = rate[1::50]
    g = GpuVectorIndex(8)
    for bias in (0.15, 1.0, 0.0):
        got, order = g.apply_score_decay(raw, idle, acc, rate, recency_bias=bias, seg_len=seg, rerank=True)
        exp = np.array([apply_score_decay(float(raw[i]), int(idle[i]), int(acc[i]), float(rate[i]), recency_bias=bias)
                        for i in range(n)], np.float32)
        # f64 exp() may differ in its last bit between libm and the device: at most one f32 ulp, and only rarely
        ulp = np.abs(got.view(np.int32).astype(np.int64) - exp.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, ulp.max()
        assert (ulp > 0).mean() < 0.001
        for s0 in range(0, n, seg):
            sc = got[s0:s0 + seg]
            ref = sorted(range(len(sc)), key=lambda i: -sc[i])  # stable descending (routes.rs:946)
            assert order[s0:s0 + seg].tolist() == ref
    off = g.apply_score_decay(raw, idle, acc, rate, enabled=False)
    assert same_bits(off, raw)
