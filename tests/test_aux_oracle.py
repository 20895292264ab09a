"""CPU: the restatements behind the extractor and the score decay (oracle/aux_oracle.c) against the
reference's own fixtures -- the golden `Node` bytes (storage/redb_storage.rs:1834-1856) and the assertions
of vector/scoring.rs:136-260."""
import os

import numpy as np

from oracle.binding import apply_score_decay, walk_node

from _nodes import canonical_node, node_bytes

GOLD = os.path.join(os.path.dirname(__file__), "golden", "node_golden.bin")


def test_builder_reproduces_the_reference_golden_bytes():
    assert canonical_node() == open(GOLD, "rb").read()


def test_walk_golden_node():
    st, nid, row, created, last, acc = walk_node(open(GOLD, "rb").read(), 384)
    assert st == 1  # embedding: None
    assert nid == bytes([0x01, 0x92, 0xab, 0xcd, 0xef, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
    assert created == 1_700_000_000 * 10**9 and last == 0 and acc == 0


def test_walk_variants():
    nid = bytes(range(16))
    e = np.arange(8, dtype=np.float32) / 7
    v = node_bytes(nid, embedding=e, tags=("x", "yy"), session="s", channel=None, access_count=42,
                   created="2024-02-29T12:00:00.250Z", last_accessed="2024-03-01T01:02:03+01:00")
    st, i, row, created, last, acc = walk_node(v, 8)
    assert st == 0 and i == nid and np.array_equal(row, e) and acc == 42
    assert created == (1709208000 * 10**9 + 250_000_000) and last == 1709251323 * 10**9
    assert walk_node(v, 16)[0] == 3                                   # dimension mismatch
    assert walk_node(node_bytes(nid, embedding=e, deleted=True), 8)[0] == 2
    meta = (1).to_bytes(8, "little") + (1).to_bytes(8, "little") + b"k" + (1).to_bytes(8, "little") + b"v"
    assert walk_node(node_bytes(nid, embedding=e, metadata_raw=meta), 8)[0] == 4
    assert walk_node(node_bytes(nid, embedding=e, created="yesterday at noon....."), 8)[0] == 4
    assert walk_node(v[:-30], 8)[0] == 5                                # truncated
    assert walk_node(node_bytes(nid, embedding=e, title=b"\xff\xfe"), 8)[0] == 5   # not UTF-8
    assert walk_node(v + b"trailing", 8)[0] == 0                         # bincode::deserialize allows trailing bytes


def test_score_decay_reference_assertions():
    """vector/scoring.rs:136-260"""
    day = 86400
    assert apply_score_decay(0.8, 0, 0, 0.01, enabled=False) == np.float32(0.8)           # :137-144
    assert apply_score_decay(0.8, 0, 0, 0.01, recency_bias=0.0) == np.float32(0.8)        # :147-151
    assert abs(apply_score_decay(0.8, 0, 0, 0.01) - 0.8) < 0.01                           # fresh node :154-166
    assert apply_score_decay(0.8, 100 * day, 0, 0.01) < 0.8                               # stale :169-185
    floor = 0.8 * (1 - 0.15) + 0.8 * 0.1 * 1.0 * 0.15
    assert abs(apply_score_decay(0.8, 400 * day, 0, 0.05) - floor) < 0.01                 # floored :188-207
    capped = 0.8 * (1 - 0.15) + 0.8 * 1.0 * 2.0 * 0.15
    assert abs(apply_score_decay(0.8, 0, 10_000, 0.01) - capped) < 0.01                   # echo cap :210-226
    assert apply_score_decay(0.8, 30 * day, 0, 0.05) < apply_score_decay(0.8, 30 * day, 0, 0.005)  # kind rates
    assert apply_score_decay(0.8, -5, 0, 0.01) == apply_score_decay(0.8, 0, 0, 0.01)      # num_seconds().max(0)
