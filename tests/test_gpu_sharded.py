"""GPU tests of the sharded path: packed-key search, key merge, the pipelined exchange and -- when the box has at least
two GPUs -- the real NCCL path (one process per GPU) against the single-index answer.
Reference contract: merge of shards == one run (retrieval/merge.result.py:22-41, gip_retrieval.py:292-306)."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden, has_cuda
from helpers import make_case

pytestmark = pytest.mark.gpu

if has_cuda():
    import torch
    from dhr_b200 import GipIndex, merge_keys, pack_keys, unpack_keys, topk_merge
    from dhr_b200.gip_retrieval import shard_bounds
    from oracle import gip_oracle as go


def _grid_case():
    g = load_golden('grouped_g6_u16_grid')
    return g, int(g['S']), int(g['G']), int(g['topk'])


def test_search_keys_equals_search():
    g, S, G, k = _grid_case()
    q = g['q_vals'].astype(np.float16)
    with GipIndex.from_arrays(g['c_vals'], g['c_idx'], n_slices=S, group=G, row_offset=1000) as ix:
        s0, r0, _ = ix.search(q, g['q_idx'], k)
        keys = torch.empty((q.shape[0], k), dtype=torch.int64, device='cuda')
        bs, nb = ix.search_keys(torch.from_numpy(q).cuda(), torch.from_numpy(g['q_idx'].astype(np.int32)).cuda(), k, keys)
        assert bs > 0 and nb == (q.shape[0] + bs - 1) // bs
        assert ix.complete() == 0
    s1, r1 = unpack_keys(keys)
    assert np.array_equal(r1, r0) and np.array_equal(s1, s0)
    assert np.array_equal(pack_keys(s0, r0), keys.cpu().numpy())


@pytest.mark.parametrize('P,k', [(3, 100), (8, 1000), (16, 1000), (2, 5000), (5, 1)])
def test_merge_keys_any_number_of_parts(P, k):
    rng = np.random.default_rng(P * 1000 + k)
    Q = 7
    scores = (rng.integers(-50, 50, size=(P, Q, k)) / 8.0).astype(np.float32)       # many exact ties
    rows = np.stack([rng.permutation(4 * P * k)[:P * k].reshape(P, k) for _ in range(Q)], axis=1).astype(np.int64)
    order = np.lexsort((rows, -scores), axis=2)                                      # each part sorted (score desc, row asc)
    scores, rows = np.take_along_axis(scores, order, 2), np.take_along_axis(rows, order, 2)
    if P > 2:
        rows[1, :, k // 2:] = -1                                                     # a short part (padding tail)
    es, er = go.merge_topk(scores, rows, k)
    keys = torch.from_numpy(pack_keys(scores, rows)).cuda()
    ms, mr, _ = merge_keys(keys)
    torch.cuda.synchronize()
    assert np.array_equal(mr.cpu().numpy(), er) and np.array_equal(ms.cpu().numpy(), es)
    if k <= 4096:                                                                     # generic (fp32, int64) merge: same answer
        gs, gr = topk_merge(scores, rows)
        assert np.array_equal(gr, er) and np.array_equal(gs, es)
        perm = rng.permutation(k)                                                     # unsorted input lists are accepted too
        gs, gr = topk_merge(scores[:, :, perm], rows[:, :, perm])
        assert np.array_equal(gr, er) and np.array_equal(gs, es)


def test_pipelined_searcher_single_rank_equals_search():
    from dhr_b200.distributed import ShardedSearcher
    case = make_case(7, 40000, 600, 32, 3, 64, 39, np.uint8, grid=True)
    k = 50
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=32, group=3) as ix:
        q = torch.from_numpy(case['q_vals'].astype(np.float16)).cuda()
        qi = torch.from_numpy(case['q_idx'].astype(np.int32)).cuda()
        s0, r0, _ = ix.search(q, qi, k)
        sr = ShardedSearcher(ix, q.shape[0], k)
        s1, r1 = sr.search(q, qi, k)
        torch.cuda.synchronize()
        assert sr.n_rerun == 0
        assert torch.equal(s0, s1) and torch.equal(r0, r1)
        s2, r2 = sr.search(q[:300], qi[:300], k)                                      # buffers are reused across calls
        torch.cuda.synchronize()
        assert torch.equal(s0[:300], s2) and torch.equal(r0[:300], r2)


def test_pipelined_searcher_overflow_rerun():
    """Adversarial row order (scores increase with the row id) overflows the candidate buffers: the keys are rewritten by
    dhr_search_complete and the exchange is redone."""
    from dhr_b200.distributed import ShardedSearcher
    n, C_ = 80000, 64
    rng = np.random.default_rng(3)
    base = rng.standard_normal((1, C_)).astype(np.float32)
    base /= np.linalg.norm(base)
    c = (np.linspace(0.1, 4.0, n, dtype=np.float32)[:, None] * base).astype(np.float16)
    q = np.repeat(base, 3, axis=0).astype(np.float16)
    k = 1000
    with GipIndex.from_arrays(c, None) as ix:
        s0, r0, _ = ix.search(q, None, k)
        assert ix.stats()['n_fallback_queries'] == 3
        sr = ShardedSearcher(ix, 3, k)
        s1, r1 = sr.search(torch.from_numpy(q).cuda(), None, k)
        torch.cuda.synchronize()
        assert sr.n_rerun == 3
        assert np.array_equal(s1.cpu().numpy(), s0) and np.array_equal(r1.cpu().numpy(), r0)


# ---- real NCCL path ------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, name, ret):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from dhr_b200.distributed import sharded_search
        g = load_golden(name)
        S, G, k = int(g['S']), int(g['G']), int(g['topk'])
        n = g['c_vals'].shape[0]
        lo, hi = shard_bounds(n, world, rank)
        q = torch.from_numpy(g['q_vals'].astype(np.float16)).cuda()
        qi = torch.from_numpy(g['q_idx'].astype(np.int32)).cuda()
        with GipIndex.from_arrays(g['c_vals'][lo:hi], g['c_idx'][lo:hi], n_slices=S, group=G, device=rank, row_offset=lo) as ix:
            s, r = sharded_search(ix, q, qi, k)
            torch.cuda.synchronize()
            s, r = s.cpu().numpy(), r.cpu().numpy()
        with GipIndex.from_arrays(g['c_vals'], g['c_idx'], n_slices=S, group=G, device=rank) as full:
            fs, fr, _ = full.search(g['q_vals'].astype(np.float16), g['q_idx'], k)
        ret[rank] = bool(np.array_equal(r, fr) and np.array_equal(s, fs) and np.array_equal(s.astype(np.float64), g['ref_scores']))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name', ['grouped_g6_u16_grid', 'delade_g1_u8_grid'])
def test_nccl_sharded_search_equals_single_index(name):
    """merged NCCL result == single-index result bit-exactly (rows, ties included) == the reference's scores."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs (run under gpurun --gpus 2)')
    import torch.multiprocessing as mp
    world = min(4, torch.cuda.device_count())
    port = _free_port()
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, name, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(ret.get(r) is True for r in range(world))
