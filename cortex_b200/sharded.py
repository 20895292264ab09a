"""Row-sharded search: one process per GPU, each holding a contiguous block of the
corpus; every rank scans its shard for the same query batch, the fixed-size local
top-k lists are exchanged with ONE all_gather (NCCL over NVLink on GPUs, gloo in
the CPU tests) and merged identically on every rank.

The reference is single-process (ARCHITECTURE.md:38); sharding is listed there as
future work (ARCHITECTURE.md:365-368).  Scores are independent per (query,row) and
top-k is mergeable, so the exchange is only B*k*16 bytes per rank.

Merge order = the single-index order: score descending, NaN last, then global row
ascending, where global row = shard offset + local row (shards are contiguous
blocks in insertion order).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

_ROW_BITS = 31
_ROW_MASK = (1 << _ROW_BITS) - 1
_INVALID = -(1 << 62)


def pack_keys(rows: torch.Tensor, score: torch.Tensor, n: torch.Tensor, row_offset: int) -> torch.Tensor:
    """(score, global row) -> sortable int64; larger = better.  rows/score [B,k], n [B]."""
    sbits = score.contiguous().view(torch.int32).to(torch.int64)
    sbits = torch.where(torch.isnan(score), torch.full_like(sbits, -1), sbits)  # NaN last
    sbits = torch.where(score == 0, torch.zeros_like(sbits), sbits)             # -0.0 ties with +0.0
    grow = rows.to(torch.int64) + int(row_offset)
    key = (sbits << _ROW_BITS) | (_ROW_MASK - grow)
    k = rows.shape[1]
    valid = torch.arange(k, device=rows.device)[None, :] < n.to(torch.int64)[:, None]
    return torch.where(valid, key, torch.full_like(key, _INVALID))


def unpack_keys(key: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (global_row int64, score_bits int32 (only for non-NaN), valid bool)"""
    valid = key > _INVALID
    grow = _ROW_MASK - (key & _ROW_MASK)
    sbits = (key >> _ROW_BITS).to(torch.int32)
    return grow, sbits, valid


def merge_gathered(keys: torch.Tensor, dists: torch.Tensor, k: int):
    """keys/dists [W,B,k] -> merged (global_rows [B,k], score [B,k], dist [B,k], n [B])."""
    W, B, kk = keys.shape
    flat = keys.permute(1, 0, 2).reshape(B, W * kk)
    dflat = dists.permute(1, 0, 2).reshape(B, W * kk)
    top, idx = torch.sort(flat, dim=1, descending=True, stable=True)
    top, idx = top[:, :k], idx[:, :k]
    grow, sbits, valid = unpack_keys(top)
    d = torch.gather(dflat, 1, idx)
    # the score is recomputed from its own bits except for NaN rows, whose distance is NaN too
    score = sbits.view(torch.float32) if sbits.is_contiguous() else sbits.contiguous().view(torch.float32)
    score = torch.where(sbits < 0, torch.full_like(score, float("nan")), score)
    n = valid.sum(dim=1).to(torch.int32)
    return grow, score, d, n


class ShardedSearch:
    """local_search(d_queries, k) -> (rows int32 [B,k], score f32 [B,k], dist f32 [B,k], n int32 [B])"""

    def __init__(self, local_search: Callable, row_offset: int, group: Optional[dist.ProcessGroup] = None,
                 local_begin: Optional[Callable] = None, local_end: Optional[Callable] = None,
                 ticket_ok_ptr: Optional[Callable] = None):
        """local_begin(d_queries, k, slot) -> (out, ticket) / local_end(ticket) -> n_redone / ticket_ok_ptr(ticket) ->
        device address: the two-step form of the local scan (GpuVectorIndex.search_batch_device_begin/_end).
        With them the exchange is enqueued behind the scan with no host wait in between."""
        self.local_search = local_search
        self.local_begin, self.local_end, self.ticket_ok_ptr = local_begin, local_end, ticket_ok_ptr
        self.row_offset = int(row_offset)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._bufs = {}

    def _exchange_cuda(self, rows, score, d, n, k, ok_ptr=0, slot=0):
        """GPU exchange: pack kernel -> one NCCL all_gather -> merge kernel (cx_merge.cu), all enqueued on
        the current stream.  Returns the merged lists and a device word = unverified queries over all ranks."""
        import ctypes as C

        from . import _capi
        L = _capi.load()
        B = rows.shape[0]
        dev = rows.device
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        key = (B, k, str(dev), slot)
        buf = self._bufs.get(key)
        if buf is None:
            words = B * k * 2 + 2  # slots + trailer (cortex_gpu.h)
            buf = (torch.empty((words,), dtype=torch.int64, device=dev),
                   torch.empty((self.world * words,), dtype=torch.int64, device=dev),
                   torch.empty((B, k), dtype=torch.int64, device=dev),
                   torch.empty((B, k), dtype=torch.float32, device=dev),
                   torch.empty((B, k), dtype=torch.float32, device=dev),
                   torch.empty((B,), dtype=torch.int32, device=dev),
                   torch.zeros((1,), dtype=torch.int64, device=dev),
                   torch.zeros((1,), dtype=torch.int64).pin_memory())
            self._bufs[key] = buf
        payload, gathered, grow, gscore, gdist, gn, unv, unv_host = buf
        st = L.cx_pack_topk_device(rows.data_ptr(), score.data_ptr(), d.data_ptr(), n.data_ptr(),
                                   C.c_void_p(ok_ptr) if ok_ptr else None, B, k, self.row_offset,
                                   payload.data_ptr(), stream)
        if st != 0:
            raise RuntimeError(L.cx_last_error().decode())
        dist.all_gather_into_tensor(gathered, payload, group=self.group)
        st = L.cx_merge_topk_device(gathered.data_ptr(), self.world, B, k, grow.data_ptr(), gscore.data_ptr(),
                                    gdist.data_ptr(), gn.data_ptr(), unv.data_ptr(), stream)
        if st != 0:
            raise RuntimeError(L.cx_last_error().decode())
        return (grow, gscore, gdist, gn), unv, unv_host

    def _search_cuda(self, rows, score, d, n, k):
        return self._exchange_cuda(rows, score, d, n, k)[0]

    def _search_overlapped(self, queries, k):
        """scan -> pack -> all_gather -> merge enqueued back to back; ONE host wait at the end.  The
        gathered trailers tell every rank whether any rank still has unverified queries; only then (rare)
        do all ranks repeat the exchange after the local retries."""
        out, ticket = self.local_begin(queries, k, 0)
        rows, score, d, n = out
        res, unv, unv_host = self._exchange_cuda(rows, score, d, n, k, self.ticket_ok_ptr(ticket))
        unv_host.copy_(unv, non_blocking=True)
        # the one host wait of the step; after it nothing reads the local lists any more, so _end may
        # overwrite them with the retried results
        torch.cuda.current_stream(rows.device).synchronize()
        self.local_end(ticket)                   # retries this rank's unverified queries, if any
        if int(unv_host.item()):                 # same value on every rank: collective decision
            res = self._exchange_cuda(rows, score, d, n, k)[0]
        return res

    # ---- pipelined form: keep the GPU fed across calls ------------------------------------------------
    def search_begin(self, queries: torch.Tensor, k: int, slot: int = 0):
        """Enqueue one sharded search (scan + exchange) and return without waiting.  `slot` selects the set
        of exchange buffers (use alternating slots to keep two searches in flight); local_begin(q, k, slot)
        must likewise write into per-slot output buffers.  Finish with search_end(pending)."""
        assert queries.is_cuda and self.local_begin is not None
        out, ticket = self.local_begin(queries, k, slot)
        if self.world == 1:
            return (out, ticket, None, None, None, k, slot)
        rows, score, d, n = out
        res, unv, unv_host = self._exchange_cuda(rows, score, d, n, k, self.ticket_ok_ptr(ticket), slot)
        unv_host.copy_(unv, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(rows.device))
        return (out, ticket, res, unv_host, ev, k, slot)

    def search_end(self, pending):
        out, ticket, res, unv_host, ev, k, slot = pending
        if self.world == 1:
            self.local_end(ticket)               # waits for this scan only; retries what was not verified
            rows, score, d, n = out
            return rows.to(torch.int64) + self.row_offset, score, d, n
        ev.synchronize()                         # this search's exchange is done; later ones keep running
        self.local_end(ticket)
        if int(unv_host.item()):                 # same value on every rank: collective decision
            rows, score, d, n = out
            res = self._exchange_cuda(rows, score, d, n, k, 0, slot)[0]
        return res

    def search(self, queries: torch.Tensor, k: int):
        if self.world > 1 and queries.is_cuda and self.local_begin is not None:
            return self._search_overlapped(queries, k)
        rows, score, d, n = self.local_search(queries, k)
        if self.world == 1:  # nothing to exchange: the local list is already the answer
            return rows.to(torch.int64) + self.row_offset, score, d, n
        if rows.is_cuda:
            return self._search_cuda(rows, score, d, n, k)
        key = pack_keys(rows, score, n, self.row_offset)
        payload = torch.stack([key, d.contiguous().view(torch.int32).to(torch.int64)], dim=0)  # [2,B,k]
        flat = torch.empty((self.world * 2,) + tuple(payload.shape[1:]), dtype=payload.dtype,
                           device=payload.device)
        dist.all_gather_into_tensor(flat, payload, group=self.group)  # concatenated along dim 0
        gathered = flat.view((self.world, 2) + tuple(payload.shape[1:]))
        keys = gathered[:, 0]
        dists = gathered[:, 1].to(torch.int32).view(torch.float32)
        return merge_gathered(keys, dists, k)

    def autolink(self, new_embeddings: torch.Tensor, self_global_rows: Optional[torch.Tensor] = None, k: int = 100,
                 threshold: float = 0.75, max_edges_per_node: int = 50):
        """The scan step of AutoLinker::run_cycle (linker/auto_linker.rs:215-264) over the sharded corpus:
        sharded search(embedding, k) (:221), skip the node itself (:235-237; `self_global_rows` [B] int64,
        -1 = not in the index), keep `score >= threshold` (linker/rules.rs:50), at most
        max_edges_per_node per node (:261).  Returns (global_rows [B,me] int64, score [B,me], n [B]);
        identical on every rank."""
        # only the first max_edges + 1 neighbours can become links (best first, self skipped at most once, the
        # walk stops below the threshold or at the cap): exchanging lists of that length is exact
        grow, score, _, n = self.search(new_embeddings, min(int(k), int(max_edges_per_node) + 1))
        return self._autolink_filter(grow, score, n, self_global_rows, threshold, max_edges_per_node, 0)

    def autolink_begin(self, new_embeddings: torch.Tensor, k: int = 100, slot: int = 0, max_edges_per_node: int = 50):
        """Pipelined form for a cycle that runs in several batches: enqueue the sharded search of this batch and
        return; autolink_end finishes it.  With alternating slots the exchange + merge of one batch overlaps
        the scan of the next."""
        return self.search_begin(new_embeddings, min(int(k), int(max_edges_per_node) + 1), slot), slot

    def autolink_end(self, pending, self_global_rows: Optional[torch.Tensor] = None, threshold: float = 0.75,
                     max_edges_per_node: int = 50):
        pend, slot = pending
        grow, score, _, n = self.search_end(pend)
        return self._autolink_filter(grow, score, n, self_global_rows, threshold, max_edges_per_node, slot)

    def _autolink_filter(self, grow, score, n, self_global_rows, threshold, max_edges_per_node, slot):
        B, k = grow.shape[0], grow.shape[1]
        me = min(int(max_edges_per_node), grow.shape[1])
        if grow.is_cuda:
            # one CUDA kernel on the merged lists (cx_merge.cu), behind the sharded search on the same stream
            import ctypes as C

            from . import _capi
            L = _capi.load()
            dev = grow.device
            key = ("al", B, k, me, str(dev), slot)
            buf = self._bufs.get(key)
            if buf is None:
                buf = (torch.empty((B, me), dtype=torch.int64, device=dev),
                       torch.empty((B, me), dtype=torch.float32, device=dev),
                       torch.empty((B,), dtype=torch.int32, device=dev))
                self._bufs[key] = buf
            out_rows, out_score, out_n = buf
            selfp = None
            if self_global_rows is not None:
                self_dev = self_global_rows.to(dev).to(torch.int64).contiguous()
                selfp = self_dev.data_ptr()
            st = L.cx_autolink_filter_device(grow.data_ptr(), score.data_ptr(), n.data_ptr(), selfp, B, grow.shape[1],
                                             C.c_float(threshold), me, out_rows.data_ptr(), out_score.data_ptr(),
                                             out_n.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if st != 0:
                raise RuntimeError(L.cx_last_error().decode())
            return out_rows, out_score, out_n
        # CPU tensors (the gloo tests of the host logic): the same rule in torch
        idx = torch.arange(grow.shape[1], device=grow.device)[None, :]
        keep = (idx < n.to(torch.int64)[:, None]) & (score >= threshold)        # NaN >= t is False
        if self_global_rows is not None:
            keep &= grow != self_global_rows.to(grow.device).to(torch.int64)[:, None]
        # stable compaction of the kept entries to the front (they are already best first)
        order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)
        out_rows = torch.gather(grow, 1, order)[:, :me]
        out_score = torch.gather(score, 1, order)[:, :me]
        out_n = keep.sum(dim=1).clamp(max=me).to(torch.int32)
        return out_rows, out_score, out_n
