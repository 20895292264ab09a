"""Seeded synthetic embeddings shared by tests and bench.

Clustered unit-norm rows (i.i.d. unit vectors at D=384 have cosine sigma ~0.05,
so nothing would ever cross the reference's 0.75 / 0.92 thresholds): draw C
centroids, row = normalise(centroid + sigma * noise).  A fraction of rows are
exact duplicates of earlier rows (ties) and, optionally, one row is all zeros
(zero norm -> NaN score in the reference arithmetic, vector/index.rs:176).

numpy's Philox bit generator is counter based, so the stream depends only on
(seed, call order); the seed is recorded in every result file.
"""
from __future__ import annotations

import numpy as np

SEED = 0xC027E5


def _rng(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=seed + (stream << 32)))


def make_ids(n: int, start: int = 0) -> np.ndarray:
    """Deterministic 16-byte ids: big-endian counter in the low 8 bytes, tag in the high."""
    ids = np.zeros((n, 16), dtype=np.uint8)
    ctr = (np.arange(start, start + n, dtype=np.uint64)).astype(">u8").view(np.uint8).reshape(n, 8)
    ids[:, 8:] = ctr
    ids[:, 0] = 0xC0
    ids[:, 1] = 0x27
    ids[:, 6] = 0x70  # looks like a v7 uuid nibble; purely cosmetic
    return ids


def make_corpus(n: int, dim: int, *, n_clusters: int | None = None, sigma: float = 0.035,
                dup_frac: float = 0.01, zero_row: bool = False, normalise: bool = True,
                seed: int = SEED, chunk: int = 1 << 17) -> np.ndarray:
    """n x dim float32 clustered rows.  sigma=0.035 at D=384 gives intra-cluster
    cosine ~ 1/(1+sigma^2*D) ~ 0.68..0.9 spread, with hits above 0.75 and 0.92."""
    if n_clusters is None:
        n_clusters = max(1, n // 64)
    rng = _rng(seed, 1)
    cent = rng.standard_normal((n_clusters, dim), dtype=np.float32)
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    out = np.empty((n, dim), dtype=np.float32)
    which = _rng(seed, 2).integers(0, n_clusters, size=n)
    # per-row noise scale varies so that some pairs exceed 0.92 and some sit near 0.75
    scale = (sigma * _rng(seed, 3).uniform(0.3, 1.6, size=n)).astype(np.float32)
    nrng = _rng(seed, 4)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = nrng.standard_normal((e - s, dim), dtype=np.float32)
        x *= scale[s:e, None]
        x += cent[which[s:e]]
        if normalise:
            x /= np.linalg.norm(x, axis=1, keepdims=True)
        else:
            x *= _rng(seed, 5 + s).uniform(0.5, 3.0, size=(e - s, 1)).astype(np.float32)
        out[s:e] = x
    n_dup = int(n * dup_frac)
    if n_dup and n > 2:
        drng = _rng(seed, 6)
        dst = drng.choice(np.arange(1, n), size=min(n_dup, n - 1), replace=False)
        src = (dst * drng.uniform(0, 1, size=dst.size)).astype(np.int64)  # an earlier row
        order = np.argsort(dst)
        for d, s_ in zip(dst[order], src[order]):
            out[d] = out[s_]
    if zero_row and n > 3:
        out[n // 3] = 0.0
    return out


def make_queries(corpus: np.ndarray, b: int, *, sigma: float = 0.02, seed: int = SEED) -> np.ndarray:
    """Half perturbed corpus rows, half fresh draws around corpus rows with larger noise."""
    n, dim = corpus.shape
    rng = _rng(seed, 100)
    pick = rng.integers(0, n, size=b)
    q = corpus[pick].astype(np.float32).copy()
    noise = rng.standard_normal((b, dim), dtype=np.float32)
    s = np.where(np.arange(b) % 2 == 0, sigma, 3.0 * sigma).astype(np.float32)
    q += noise * s[:, None]
    nrm = np.linalg.norm(q, axis=1, keepdims=True)
    nrm[nrm == 0] = 1.0
    q /= nrm
    return q
