#!/usr/bin/env python
"""Copies the reference's golden `Node` bytes (bincode 1.3 of make_canonical_node(),
/root/reference/crates/cortex-core/src/storage/redb_storage.rs:1834-1856) into
tests/golden/node_golden.bin, so that the layout the extractor walks is pinned by the reference's own
regression fixture on machines where /root/reference does not exist.  Run in the build container:
    python tests/golden/make_node_golden.py
"""
import os
import re

SRC = "/root/reference/crates/cortex-core/src/storage/redb_storage.rs"
HERE = os.path.dirname(os.path.abspath(__file__))

txt = open(SRC).read()
m = re.search(r"const GOLDEN_NODE_BYTES: &\[u8\] = &\[(.*?)\];", txt, re.S)
assert m, "golden bytes not found"
vals = [int(x) for x in re.findall(r"\d+", m.group(1))]
assert all(0 <= v < 256 for v in vals)
open(os.path.join(HERE, "node_golden.bin"), "wb").write(bytes(vals))
print(len(vals), "bytes written")
