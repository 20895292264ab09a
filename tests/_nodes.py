"""Builder of serialized `Node` values (bincode 1.3, fixint, little endian) for the extractor tests --
the byte layout of /root/reference/crates/cortex-core/src/types.rs:26-68 as pinned by the reference's
golden bytes (tests/golden/node_golden.bin reproduces from canonical_node() below)."""
import struct

import numpy as np


def _s(x) -> bytes:
    b = x if isinstance(x, bytes) else x.encode()
    return struct.pack("<Q", len(b)) + b


def _opt_s(x) -> bytes:
    return b"\x00" if x is None else b"\x01" + _s(x)


def node_bytes(node_id: bytes, kind="fact", title="t", body="b", metadata_raw: bytes = None, tags=(), embedding=None,
               agent="a", session=None, channel=None, importance=0.5, access_count=0,
               last_accessed="1970-01-01T00:00:00Z", created="2023-11-14T22:13:20Z", updated="2023-11-14T22:13:20Z",
               deleted=False) -> bytes:
    out = struct.pack("<Q", 16) + bytes(node_id)
    out += _s(kind) + _s(title) + _s(body)
    out += struct.pack("<Q", 0) if metadata_raw is None else metadata_raw
    out += struct.pack("<Q", len(tags)) + b"".join(_s(t) for t in tags)
    if embedding is None:
        out += b"\x00"
    else:
        e = np.asarray(embedding, np.float32)
        out += b"\x01" + struct.pack("<Q", e.size) + e.astype("<f4").tobytes()
    out += _s(agent) + _opt_s(session) + _opt_s(channel)
    out += struct.pack("<f", importance) + struct.pack("<Q", access_count)
    out += _s(last_accessed) + _s(created) + _s(updated)
    out += b"\x01" if deleted else b"\x00"
    return out


def canonical_node() -> bytes:
    """make_canonical_node() of storage/redb_storage.rs:1165-1194"""
    return node_bytes(bytes([0x01, 0x92, 0xab, 0xcd, 0xef, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]), kind="fact",
                      title="Schema regression test", body="This node is used to detect Node struct changes.",
                      tags=("regression",), embedding=None, agent="test-agent", importance=0.5, access_count=0,
                      last_accessed="1970-01-01T00:00:00Z", created="2023-11-14T22:13:20Z",
                      updated="2023-11-14T22:13:20Z", deleted=False)


def pack(values):
    """values -> (blob uint8, offsets uint64 [n+1])"""
    offs = np.zeros(len(values) + 1, np.uint64)
    offs[1:] = np.cumsum([len(v) for v in values])
    return np.frombuffer(b"".join(values), np.uint8).copy(), offs
