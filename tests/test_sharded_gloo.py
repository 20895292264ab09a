"""world_size-2 gloo test of the row-sharded merge (the N>1 path's host logic):
each rank scans half of the corpus with the CPU oracle standing in for the local
GPU scan, the all_gather + merge must reproduce the single-index answer exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cortex_b200 import synth
from cortex_b200.sharded import ShardedSearch, merge_gathered, pack_keys
from oracle.binding import OracleIndex


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, b, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    corpus = synth.make_corpus(n, d, zero_row=True, seed=31)
    ids = synth.make_ids(n)
    Q = synth.make_queries(corpus, b, seed=31)
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)
    local = OracleIndex(d, faithful_copy=False)
    local.insert_batch(ids[lo:hi], corpus[lo:hi])

    def local_search(q, kk):
        _, sc, di, rows, nn = local.search_batch(q.numpy(), kk)
        return (torch.from_numpy(rows.astype(np.int32)), torch.from_numpy(sc), torch.from_numpy(di),
                torch.from_numpy(nn.astype(np.int32)))

    sh = ShardedSearch(local_search, row_offset=lo)
    grow, score, dd, nn = sh.search(torch.from_numpy(Q), k)
    # auto-link step over the shards: the first b corpus rows play the new nodes (they are in the index)
    self_rows = torch.arange(b, dtype=torch.int64)
    arow, asc, an = sh.autolink(torch.from_numpy(corpus[:b].copy()), self_rows, k=20, threshold=0.75,
                                max_edges_per_node=5)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), grow=grow.numpy(), score=score.numpy(), dist=dd.numpy(),
             n=nn.numpy(), arow=arow.numpy(), asc=asc.numpy(), an=an.numpy())
    dist.destroy_process_group()


def test_two_rank_merge_matches_single_index(tmp_path):
    n, d, b, k = 3000, 64, 7, 10
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, d, b, k, str(tmp_path)), nprocs=2, join=True)
    corpus = synth.make_corpus(n, d, zero_row=True, seed=31)
    Q = synth.make_queries(corpus, b, seed=31)
    full = OracleIndex(d, faithful_copy=False)
    full.insert_batch(synth.make_ids(n), corpus)
    _, sc, di, rows, nn = full.search_batch(Q, k)
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert np.array_equal(z["n"], nn.astype(np.int32))
        assert np.array_equal(z["grow"], rows.astype(np.int64))
        assert np.array_equal(z["score"].view(np.uint32), sc.view(np.uint32))
        assert np.array_equal(z["dist"].view(np.uint32), di.view(np.uint32))
    # auto-link: the reference loop (auto_linker.rs:215-264) on the single index
    _, sc20, _, rows20, n20 = full.search_batch(corpus[:b], 20)
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        for q in range(b):
            exp = [(int(rows20[q, j]), sc20[q, j]) for j in range(int(n20[q]))
                   if int(rows20[q, j]) != q and sc20[q, j] >= np.float32(0.75)][:5]
            got = [(int(z["arow"][q, j]), z["asc"][q, j]) for j in range(int(z["an"][q]))]
            assert [g[0] for g in got] == [e[0] for e in exp], (q, got, exp)
            assert all(np.float32(g[1]).view(np.uint32) == np.float32(e[1]).view(np.uint32) for g, e in zip(got, exp))


def test_merge_orders_ties_nan_and_short_lists():
    # rank 0 rows 0..2, rank 1 rows 10..12 ; equal scores tie on global row; NaN last; short list padded
    s0 = torch.tensor([[0.9, 0.5, float("nan")]])
    s1 = torch.tensor([[0.9, 0.5, 0.0]])
    r0 = torch.tensor([[2, 0, 1]], dtype=torch.int32)
    r1 = torch.tensor([[0, 1, 2]], dtype=torch.int32)
    k0 = pack_keys(r0, s0, torch.tensor([3], dtype=torch.int32), 0)
    k1 = pack_keys(r1, s1, torch.tensor([2], dtype=torch.int32), 10)
    d0, d1 = 1 - s0, 1 - s1
    grow, score, d, n = merge_gathered(torch.stack([k0, k1]), torch.stack([d0, d1]), 6)
    assert n.tolist() == [5]
    assert grow[0, :5].tolist() == [2, 10, 0, 11, 1]
    assert score[0, :4].tolist() == [0.9, 0.9, 0.5, 0.5] or np.allclose(score[0, :4], [0.9, 0.9, 0.5, 0.5])
    assert torch.isnan(score[0, 4])
