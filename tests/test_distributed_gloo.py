"""world_size-2 gloo test (CPU) of the multi-GPU host logic: range sharding + all-gather plumbing + merge.
The per-shard search results and the merge are supplied by the oracle here (test doubles): the CUDA search and the
merge kernel themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from oracle import gip_oracle as go


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleShard:
    """Stands in for GipIndex on a CPU box: exact per-shard top-k with global rows."""

    def __init__(self, g, lo, hi):
        self.g, self.lo, self.hi, self.row_offset = g, lo, hi, lo

    def search(self, q_vals, q_idx, k, lamda=1.0, masked=True, out=None, return_torch=True):
        g = self.g
        rows, vals = go.search_f64(q_vals, q_idx, g['c_vals'][self.lo:self.hi], g['c_idx'][self.lo:self.hi],
                                   int(g['S']), int(g['G']), k)
        return torch.from_numpy(vals.astype(np.float32)), torch.from_numpy(rows + self.lo), None


def _oracle_merge(gs, gr):
    s, r = go.merge_topk(gs.numpy(), gr.numpy(), gs.shape[2])
    return torch.from_numpy(s), torch.from_numpy(r)


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dhr_b200.distributed import local_shard, sharded_search_lists
        g = load_golden('delade_g1_u8_grid')
        n = g['c_vals'].shape[0]
        lo, hi = local_shard(n)
        assert (lo, hi) == go.shard_bounds(n, world, rank)
        q = g['q_vals'].astype(np.float32)
        k = int(g['topk'])
        shard = _OracleShard(g, lo, hi)
        s, r = sharded_search_lists(lambda: shard.search(q, g['q_idx'], k)[:2], merge_fn=_oracle_merge)
        full_r, full_s = go.search_f64(q, g['q_idx'], g['c_vals'], g['c_idx'], int(g['S']), int(g['G']), k)
        ok = np.array_equal(r.numpy(), full_r) and np.array_equal(s.numpy(), full_s.astype(np.float32))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single_shard():
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True
