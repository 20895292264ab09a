"""The CUDA exchange kernels of the row-sharded search (cx_merge.cu) against the torch
reference merge that the world_size-2 gloo test validates against the single-index
oracle, plus a simulated 4-shard search on one GPU compared with the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from cortex_b200 import GpuVectorIndex, _capi, synth
from cortex_b200.sharded import merge_gathered, pack_keys
from oracle.binding import OracleIndex

pytestmark = pytest.mark.gpu


def cuda_pack(L, rows, score, dist, n, offset, ok=None):
    B, k = rows.shape
    payload = torch.empty((B * k * 2 + 2,), dtype=torch.int64, device=rows.device)  # slots + trailer
    st = L.cx_pack_topk_device(rows.data_ptr(), score.data_ptr(), dist.data_ptr(), n.data_ptr(),
                               ok.data_ptr() if ok is not None else None, B, k, offset,
                               payload.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == 0, L.cx_last_error()
    return payload


def cuda_merge(L, gathered, k, B, want_unverified=False):
    W = gathered.shape[0]
    dev = gathered.device
    unv = torch.full((1,), -1, dtype=torch.int64, device=dev)
    grow = torch.empty((B, k), dtype=torch.int64, device=dev)
    gs = torch.empty((B, k), dtype=torch.float32, device=dev)
    gd = torch.empty((B, k), dtype=torch.float32, device=dev)
    gn = torch.empty((B,), dtype=torch.int32, device=dev)
    st = L.cx_merge_topk_device(gathered.data_ptr(), W, B, k, grow.data_ptr(), gs.data_ptr(), gd.data_ptr(),
                                gn.data_ptr(), unv.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == 0, L.cx_last_error()
    torch.cuda.synchronize()
    if want_unverified:
        return grow, gs, gd, gn, int(unv.item())
    return grow, gs, gd, gn


def test_four_simulated_shards_match_single_index_oracle():
    L = _capi.load()
    n, d, b, k, W = 8000, 384, 33, 10, 4
    corpus = synth.make_corpus(n, d, zero_row=True, dup_frac=0.05, seed=77)
    ids = synth.make_ids(n)
    Q = synth.make_queries(corpus, b, seed=77)
    per = n // W
    payloads = []
    dq = torch.from_numpy(Q).cuda()
    for w in range(W):
        g = GpuVectorIndex(d)
        g.insert_batch(ids[w * per:(w + 1) * per], corpus[w * per:(w + 1) * per])
        rows, sc, di, nn = g.search_batch_device(dq, k, stream=torch.cuda.current_stream().cuda_stream)
        payloads.append(cuda_pack(L, rows, sc, di, nn, w * per))
    grow, gs, gd, gn = cuda_merge(L, torch.stack(payloads).contiguous(), k, b)
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(ids, corpus)
    _, osc, odi, orow, on = o.search_batch(Q, k)
    assert np.array_equal(gn.cpu().numpy(), on.astype(np.int32))
    assert np.array_equal(grow.cpu().numpy(), orow.astype(np.int64))
    a, bb = gs.cpu().numpy(), osc
    assert np.all((a.view(np.uint32) == bb.view(np.uint32)) | (np.isnan(a) & np.isnan(bb)))
    a, bb = gd.cpu().numpy(), odi
    assert np.all((a.view(np.uint32) == bb.view(np.uint32)) | (np.isnan(a) & np.isnan(bb)))


def test_cuda_merge_equals_torch_merge_with_ties_nan_and_short_lists():
    L = _capi.load()
    torch.manual_seed(5)
    W, B, k = 3, 40, 7
    payloads, keys, dists = [], [], []
    for w in range(W):
        dist = (torch.randint(0, 6, (B, k)).float() * 0.25)  # few distinct values -> many ties
        dist[torch.rand(B, k) < 0.1] = float("nan")
        dist, _ = torch.sort(dist, dim=1)
        score = (1.0 - dist).clamp(0, 1)
        # a rank's list is in merged order: score descending, NaN last, equal scores by ascending row
        rows = torch.stack([torch.sort(torch.randperm(1000)[:k]).values for _ in range(B)]).int()
        n = torch.randint(0, k + 1, (B,)).int()
        r, s, dd, nn = rows.cuda(), score.cuda(), dist.cuda(), n.cuda()
        payloads.append(cuda_pack(L, r, s, dd, nn, w * 1000))
        keys.append(pack_keys(rows, score, n, w * 1000))
        dists.append(dist)
    grow, gs, gd, gn = cuda_merge(L, torch.stack(payloads).contiguous(), k, B)
    trow, ts, td, tn = merge_gathered(torch.stack(keys), torch.stack(dists), k)
    assert torch.equal(gn.cpu(), tn)
    for b in range(B):
        m = int(tn[b])
        assert torch.equal(grow.cpu()[b, :m], trow[b, :m])
        x, y = gs.cpu()[b, :m], ts[b, :m]
        assert torch.all((x == y) | (torch.isnan(x) & torch.isnan(y)))


def test_begin_end_and_the_unverified_trailer():
    """cx_search_batch_device_begin/_end: the exchange is enqueued behind the scan without a host wait;
    the pack trailer carries each rank's count of unverified queries so that all ranks learn from the
    gathered payloads whether the exchange must be repeated.  A zero query cannot be verified by the
    fast passes (its norm is degenerate) and forces that path."""
    L = _capi.load()
    n, d, b, k = 6000, 384, 40, 10
    corpus = synth.make_corpus(n, d, seed=12)
    Q = synth.make_queries(corpus, b, seed=13)
    Q[7] = 0.0   # degenerate norm: never verified by a fast pass
    Q[21] = 0.0
    g = GpuVectorIndex(d)
    g.insert_batch(synth.make_ids(n), corpus)
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(synth.make_ids(n), corpus)
    dq = torch.from_numpy(Q).cuda()
    out, ticket = g.search_batch_device_begin(dq, k, stream=torch.cuda.current_stream().cuda_stream)
    ok = g.ticket_ok(ticket, b)
    p = cuda_pack(L, out[0], out[1], out[2], out[3], 0, ok)
    _, _, _, _, unv = cuda_merge(L, p[None, :].contiguous(), k, b, want_unverified=True)
    assert unv == 2
    redone = g.search_batch_device_end(ticket)
    assert redone == 2
    # after _end the buffers hold the final answer
    _, osc, odi, orow, on = o.search_batch(Q, k)
    assert np.array_equal(out[3].cpu().numpy(), on.astype(np.int32))
    assert np.array_equal(out[0].cpu().numpy().astype(np.int64), orow.astype(np.int64))
    a = out[1].cpu().numpy()
    assert np.all((a.view(np.uint32) == osc.view(np.uint32)) | (np.isnan(a) & np.isnan(osc)))
    # nothing unverified: the trailer is zero and _end reports nothing to redo
    Q2 = synth.make_queries(corpus, b, seed=14)
    dq2 = torch.from_numpy(Q2).cuda()
    out2, t2 = g.search_batch_device_begin(dq2, k, stream=torch.cuda.current_stream().cuda_stream)
    p2 = cuda_pack(L, out2[0], out2[1], out2[2], out2[3], 0, g.ticket_ok(t2, b))
    _, _, _, _, unv2 = cuda_merge(L, p2[None, :].contiguous(), k, b, want_unverified=True)
    assert unv2 == 0 and g.search_batch_device_end(t2) == 0


def test_pipelined_search_keeps_two_batches_in_flight():
    """ShardedSearch.search_begin / search_end (world 1): the next batch is enqueued before the host waits
    for the previous one; every batch still comes back exact."""
    from cortex_b200.sharded import ShardedSearch

    n, d, b, k = 9000, 384, 64, 10
    corpus = synth.make_corpus(n, d, seed=31)
    g = GpuVectorIndex(d)
    g.insert_batch(synth.make_ids(n), corpus)
    o = OracleIndex(d, faithful_copy=False)
    o.insert_batch(synth.make_ids(n), corpus)
    stream = torch.cuda.current_stream().cuda_stream
    outs = {}

    def local_begin(q, kk, slot=0):
        out, ticket = g.search_batch_device_begin(q, kk, stream=stream, out=outs.get(slot))
        outs[slot] = out
        return out, ticket

    sh = ShardedSearch(lambda q, kk: g.search_batch_device(q, kk, stream=stream), row_offset=0,
                       local_begin=local_begin, local_end=g.search_batch_device_end, ticket_ok_ptr=g.ticket_ok_ptr)
    batches = [synth.make_queries(corpus, b, seed=40 + i) for i in range(5)]
    batches[2][3] = 0.0  # one query that only the exact path can answer
    dqs = [torch.from_numpy(q).cuda() for q in batches]
    pend, got = [], []
    for i, dq in enumerate(dqs):
        pend.append((i, sh.search_begin(dq, k, slot=i % 2)))
        if len(pend) == 2:
            j, p = pend.pop(0)
            rows, sc, di, nn = sh.search_end(p)
            got.append((j, rows.cpu().numpy().copy(), sc.cpu().numpy().copy(), nn.cpu().numpy().copy()))
    while pend:
        j, p = pend.pop(0)
        rows, sc, di, nn = sh.search_end(p)
        got.append((j, rows.cpu().numpy().copy(), sc.cpu().numpy().copy(), nn.cpu().numpy().copy()))
    assert [j for j, *_ in got] == list(range(5))
    for j, rows, sc, nn in got:
        _, osc, _, orow, on = o.search_batch(batches[j], k)
        assert np.array_equal(nn, on.astype(np.int32))
        assert np.array_equal(rows, orow.astype(np.int64))
        assert np.all((sc.view(np.uint32) == osc.view(np.uint32)) | (np.isnan(sc) & np.isnan(osc)))
