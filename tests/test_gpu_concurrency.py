"""search* are re-entrant (include/cortex_gpu.h): the reference shares the index as
Arc<RwLock<_>> and runs searches from many handler threads under read guards
(serve.rs:101, http/routes.rs:906, grpc/service.rs:667).  Eight host threads search one
index at once through every pass; each result must equal the oracle's."""
import threading

import numpy as np
import pytest

from cortex_b200 import GpuVectorIndex, synth
from oracle.binding import OracleIndex

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def test_concurrent_searches_from_many_threads():
    n, d = 20_000, 384
    corpus = synth.make_corpus(n, d, seed=321)
    ids = synth.make_ids(n)
    g = GpuVectorIndex(d)
    g.insert_batch(ids, corpus)
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(ids, corpus)
    # one job per pass: streaming (B=1, 3), tensor (B=40, 200), threshold on both
    jobs = []
    for t, b in enumerate((1, 3, 40, 200, 2, 64, 1, 130)):
        Q = synth.make_queries(corpus, b, seed=1000 + t)
        if t in (4, 5):
            exp = [o.search_threshold(q, 0.75) for q in Q]
            jobs.append(("thr", Q, exp))
        else:
            jobs.append(("topk", Q, o.search_batch(Q, 10)))
    errors = []

    def work(kind, Q, exp):
        try:
            for _ in range(6):
                if kind == "thr":
                    gi, gs, gd, gn, gt = g.search_threshold_batch_arrays(Q, 0.75, 128)
                    for b in range(Q.shape[0]):
                        m = min(128, len(exp[b].ids))
                        assert int(gt[b]) == len(exp[b].ids) and int(gn[b]) == m
                        assert np.array_equal(gi[b, :m], exp[b].ids[:m]) and same_bits(gs[b, :m], exp[b].score[:m])
                else:
                    oi, osc, od, _, on = exp
                    gi, gs, gd, gn = g.search_batch_arrays(Q, 10)
                    assert np.array_equal(gn, on) and np.array_equal(gi, oi)
                    assert same_bits(gs, osc) and same_bits(gd, od)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
