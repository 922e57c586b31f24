"""Multi-GPU search: corpus range-sharded across ranks (one process per GPU), per-shard top-k, one all-gather,
device-side merge.  Replaces the process-per-shard + text-file merge of the reference
(gip_retrieval.py:292-306 `--total_shrad/--shrad`, retrieval/merge.result.py:20-43).

The exchange step is a single `all_gather` of the per-shard [Q, k] (fp32 score, int64 GLOBAL row) lists over
NCCL / NVLink; every rank then holds [P, Q, k] and runs the merge kernel (dhr_topk_merge), so all ranks return
the same answer.  Exactness: each shard orders by (score desc, global row asc), so the merged list equals the
single-shard result including ties.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .gip_retrieval import shard_bounds
from .index import topk_merge


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


def sharded_search(index, q_vals, q_idx, k, lamda=1.0, masked=True, group=None, merge_fn=topk_merge, local_out=None,
                   gather_out=None):
    """Search this rank's shard (index.row_offset = first global row) and merge across ranks.
    Returns (scores [Q,k], rows [Q,k] global ids), identical on every rank."""
    scores, rows, _ = index.search(q_vals, q_idx, k, lamda=lamda, masked=masked, out=local_out, return_torch=True)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return scores, rows
    gs, gr = gather_topk(scores, rows, group, gather_out)
    return merge_fn(gs, gr)
