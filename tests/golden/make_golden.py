"""Generates tests/golden/scan_golden.npz with an INDEPENDENT numpy restatement
of the reference scan (not the C oracle), so that two separately written
restatements have to agree bit for bit.

The reference is Rust and cannot be run in this image (no cargo/rustc), so
these are not outputs of the reference itself; they follow
/root/reference/crates/cortex-core/src/vector/index.rs:169-179 (distance),
:253-256 (distance_to_similarity) and :259-294 (brute-force scan) with
np.float32 scalars, one rounding per operation, strictly left to right.

Run:  python tests/golden/make_golden.py   (rewrites the .npz next to it)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from cortex_b200 import synth  # noqa: E402

f32 = np.float32


def seq_dot(a, b):
    acc = f32(0.0)
    for x, y in zip(a, b):  # zip truncates like Rust's Iterator::zip
        acc = f32(acc + f32(f32(x) * f32(y)))
    return acc


def distance(a, b):
    dot = seq_dot(a, b)
    na = f32(np.sqrt(seq_dot(a, a)))
    nb = f32(np.sqrt(seq_dot(b, b)))
    with np.errstate(all="ignore"):
        sim = f32(dot / f32(na * nb))
    return f32(f32(1.0) - sim)


def to_similarity(d):
    s = f32(f32(1.0) - d)
    if s < f32(0.0):
        s = f32(0.0)
    if s > f32(1.0):
        s = f32(1.0)
    return s


def scan(corpus, q):
    """All rows: (row, score, distance) sorted by score desc, NaN last, ties in row order."""
    hits = []
    for r in range(corpus.shape[0]):
        d = distance(q, corpus[r])
        hits.append((r, to_similarity(d), d))
    good = [h for h in hits if not np.isnan(h[1])]
    bad = [h for h in hits if np.isnan(h[1])]
    good.sort(key=lambda h: -float(h[1]))  # python sort is stable
    return good + bad


def main():
    out = {}
    # case A: the reference's own 3-d unit-test vectors (index.rs:484-510, 687-708)
    A = np.array([[1.0, 0.0, 0.0], [0.9, 0.1, 0.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0],
                  [0.0, 0.0, 1.0]], dtype=f32)
    qa = np.array([1.0, 0.0, 0.0], dtype=f32)
    h = scan(A, qa)
    out["A_corpus"], out["A_query"] = A, qa
    out["A_rows"] = np.array([x[0] for x in h], np.uint32)
    out["A_score"] = np.array([x[1] for x in h], f32)
    out["A_dist"] = np.array([x[2] for x in h], f32)

    # case B: 96 clustered rows x 384-d with duplicates and a zero row, 6 queries
    B = synth.make_corpus(96, 384, n_clusters=4, dup_frac=0.05, zero_row=True, seed=synth.SEED + 1)
    qb = synth.make_queries(B, 6, seed=synth.SEED + 1)
    out["B_corpus"], out["B_query"] = B, qb
    rows, sc, di = [], [], []
    for q in qb:
        h = scan(B, q)
        rows.append([x[0] for x in h])
        sc.append([x[1] for x in h])
        di.append([x[2] for x in h])
    out["B_rows"] = np.array(rows, np.uint32)
    out["B_score"] = np.array(sc, f32)
    out["B_dist"] = np.array(di, f32)

    # case C: non-normalised 40 x 100-d (D not a multiple of 4 or 32), 3 queries
    Cc = synth.make_corpus(40, 100, n_clusters=3, normalise=False, dup_frac=0.0, seed=synth.SEED + 2)
    qc = (synth.make_queries(Cc, 3, seed=synth.SEED + 2) * f32(2.5)).astype(f32)
    out["C_corpus"], out["C_query"] = Cc, qc
    rows, sc, di = [], [], []
    for q in qc:
        h = scan(Cc, q)
        rows.append([x[0] for x in h])
        sc.append([x[1] for x in h])
        di.append([x[2] for x in h])
    out["C_rows"] = np.array(rows, np.uint32)
    out["C_score"] = np.array(sc, f32)
    out["C_dist"] = np.array(di, f32)

    # case D: mismatched query length (zip truncation, index.rs:172): query of 5 vs rows of 8
    D = synth.make_corpus(12, 8, n_clusters=2, dup_frac=0.0, seed=synth.SEED + 3)
    qd = np.array([0.3, -0.2, 0.9, 0.1, 0.05], dtype=f32)
    h = scan(D, qd)
    out["D_corpus"], out["D_query"] = D, qd
    out["D_rows"] = np.array([x[0] for x in h], np.uint32)
    out["D_score"] = np.array([x[1] for x in h], f32)
    out["D_dist"] = np.array([x[2] for x in h], f32)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scan_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
