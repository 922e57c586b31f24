"""Multi-GPU search: corpus range-sharded across ranks (one process per GPU), per-shard top-k, NCCL all-gather,
device-side merge.  Replaces the process-per-shard + text-file merge of the reference
(gip_retrieval.py:292-306 `--total_shrad/--shrad`, retrieval/merge.result.py:20-43).

CUDA path (`sharded_search` on an NCCL group): the shard search is ENQUEUED as one stream-ordered call
(dhr_search_keys) that emits, per batch of 256 queries, packed 64-bit keys (score bits << 32 | 0xFFFFFFFF - GLOBAL row);
a side stream waits for batch b, all-gathers its [B, k] keys (one collective of 8-byte items instead of an fp32 and an
int64 one) and merges the P sorted lists (dhr_merge_keys) while the main stream is already scanning batch b+1, so the
exchange is hidden behind the scan.  Exactness: every shard orders by (score desc, global row asc) = descending key
order, so the merged list equals the single-shard result including ties.

Host / gloo path (`sharded_search_lists`): one all-gather of ([Q,k] fp32, [Q,k] int64) + dhr_topk_merge; used by the CPU
tests of the host logic and by callers that already hold per-shard lists.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .gip_retrieval import shard_bounds
from .index import merge_keys, topk_merge


def local_shard(n_docs, group=None):
    """Rows [lo, hi) owned by this rank under the reference's rule (floor(N/P) each, remainder to the last)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    return shard_bounds(n_docs, world, rank)


def gather_topk(scores, rows, group=None, out=None):
    """all-gather per-shard [Q, k] lists into [P, Q, k] (scores fp32, rows int64) on every rank."""
    world = dist.get_world_size(group)
    Q, k = scores.shape
    if out is None:
        out = (torch.empty((world, Q, k), dtype=scores.dtype, device=scores.device),
               torch.empty((world, Q, k), dtype=rows.dtype, device=rows.device))
    gs, gr = out
    if scores.is_cuda:
        dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)
        dist.all_gather_into_tensor(gr, rows.contiguous(), group=group)
    else:   # gloo
        dist.all_gather(list(gs.unbind(0)), scores.contiguous(), group=group)
        dist.all_gather(list(gr.unbind(0)), rows.contiguous(), group=group)
    return gs, gr


def sharded_search_lists(search_fn, merge_fn=topk_merge, group=None, gather_out=None):
    """Generic form: search_fn() -> (scores [Q,k], rows [Q,k] GLOBAL ids) of this rank's shard; one all-gather + merge."""
    scores, rows = search_fn()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return scores, rows
    gs, gr = gather_topk(scores, rows, group, gather_out)
    return merge_fn(gs, gr)


class ShardedSearcher:
    """Pipelined shard search + exchange for one index (one per rank).  Buffers are allocated once and reused."""

    def __init__(self, index, n_queries, k, group=None):
        self.index, self.group = index, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = torch.device('cuda', index.device)
        self.keys = torch.empty((n_queries, k), dtype=torch.int64, device=dev)
        self.side = torch.cuda.Stream(device=dev)
        self.gathered = None            # [P, B, k], allocated when the batch size is known
        self.out = (torch.empty((n_queries, k), dtype=torch.float32, device=dev),
                    torch.empty((n_queries, k), dtype=torch.int64, device=dev))
        self.breakdown = {}
        self.profile = False            # bench.py: time the side-stream exchange (CUDA events) into self.breakdown
        self.n_merge_launches = 0
        self.n_rerun = 0

    def search(self, q_vals, q_idx, k, lamda=1.0, masked=True, out=None):
        """Returns (scores [Q,k], rows [Q,k] global ids), identical on every rank.  The result tensors are valid once the
        CURRENT stream has passed the point where this call returns (the side stream is joined back into it)."""
        ix, world = self.index, self.world
        out_s, out_r = out if out is not None else self.out
        n = q_vals.shape[0]
        main = torch.cuda.current_stream(self.keys.device)
        keys = self.keys[:n]
        bs, nb = ix.search_keys(q_vals, q_idx, k, keys, lamda=lamda, masked=masked, stream=main.cuda_stream)
        if self.gathered is None or self.gathered.shape[1] < bs:
            self.gathered = torch.empty((world, bs, k), dtype=torch.int64, device=keys.device)
        side = self.side
        prof = self.profile
        if prof:
            ev_main_end = torch.cuda.Event(enable_timing=True)
            ev_main_end.record(main)                                     # after the last scan / select launch of the call
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nb)]
        with torch.cuda.stream(side):
            for b in range(nb):
                lo, hi = b * bs, min(n, (b + 1) * bs)
                ix.wait_batch(b, side.cuda_stream)
                if prof:
                    evs[b][0].record(side)
                if world == 1:
                    merge_keys(keys[lo:hi].unsqueeze(0), out_s[lo:hi], out_r[lo:hi], stream=side.cuda_stream)
                else:
                    g = self.gathered.view(-1)[:world * (hi - lo) * k].view(world, hi - lo, k)
                    dist.all_gather_into_tensor(g, keys[lo:hi], group=self.group)
                    merge_keys(g, out_s[lo:hi], out_r[lo:hi], stream=side.cuda_stream)
                if prof:
                    evs[b][1].record(side)
        self.n_merge_launches = nb
        n_rerun = ix.complete(stream=main.cuda_stream)
        if prof:
            side.synchronize()
            self.breakdown = {'exchange_ms': sum(a.elapsed_time(b_) for a, b_ in evs), 'tail_ms': max(0.0, ev_main_end.elapsed_time(evs[-1][1])),
                              'batches': nb}
        main.wait_stream(side)
        if n_rerun > 0:
            # adversarial row order (never on shuffled corpora): some keys were rewritten after their batch was exchanged;
            # redo the exchange in one piece
            if world == 1:
                merge_keys(keys.unsqueeze(0), out_s[:n], out_r[:n], stream=main.cuda_stream)
            else:
                g = torch.empty((world, n, k), dtype=torch.int64, device=keys.device)
                dist.all_gather_into_tensor(g, keys, group=self.group)
                merge_keys(g, out_s[:n], out_r[:n], stream=main.cuda_stream)
        self.n_rerun = n_rerun
        return out_s[:n], out_r[:n]


def sharded_search(index, q_vals, q_idx, k, lamda=1.0, masked=True, group=None, searcher=None, out=None):
    """Search this rank's shard (index.row_offset = first global row) and merge across ranks (pipelined exchange).
    Returns (scores [Q,k], rows [Q,k] global ids), identical on every rank."""
    if searcher is None:
        searcher = ShardedSearcher(index, q_vals.shape[0], k, group)
    return searcher.search(q_vals, q_idx, k, lamda=lamda, masked=masked, out=out)
