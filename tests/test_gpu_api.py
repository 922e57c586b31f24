"""GPU tests of the reference-facing API: the Python mirror of gip_retrieval.py (functions, CLI, files)."""
import os
import pickle

import numpy as np
import pytest

from conftest import load_golden, has_cuda
from helpers import tie_groups_equal, make_case, assert_matches_oracle
from oracle import gip_oracle as go

pytestmark = pytest.mark.gpu

if has_cuda():
    import torch
    import dhr_b200
    from dhr_b200 import gip_retrieval as gr
    from dhr_b200 import GipIndex


def _q(g, lam=1.0):
    q = g['q_vals'].astype(np.float32)
    C = int(g['C'])
    if C > 0:
        q[:, -C:] = np.float32(lam) * q[:, -C:]
    return q


def test_gip_retrieval_function_matches_reference_exact_branch():
    g = load_golden('delade_g1_u8_grid')
    S, k = int(g['S']), int(g['topk'])
    qids = list(range(100, 100 + g['q_vals'].shape[0]))
    args = go.make_args(emb_dim=S, topk=k, brute_force=True, theta=0.3)
    res, sc = gr.GIP_retrieval(qids, torch.from_numpy(_q(g)), torch.from_numpy(g['q_idx']),
                               torch.from_numpy(g['c_vals'].astype(np.float32)), torch.from_numpy(g['c_idx']), args)
    assert args.theta == 0                                        # mutated like gip_retrieval.py:90
    assert list(res.keys()) == qids and isinstance(res[qids[0]], list) and isinstance(sc[qids[0]][0], float)
    for i, qid in enumerate(qids):
        assert np.array_equal(np.array(sc[qid]), g['ref_scores'][i])
        assert tie_groups_equal(np.array(res[qid]), g['ref_rows'][i], np.array(sc[qid]))
    with pytest.raises(RuntimeError):                             # torch.topk raises for k > N (:123)
        gr.GIP_retrieval(qids, torch.from_numpy(_q(g)), torch.from_numpy(g['q_idx']),
                         torch.from_numpy(g['c_vals'].astype(np.float32)), torch.from_numpy(g['c_idx']),
                         go.make_args(emb_dim=S, topk=g['c_vals'].shape[0] + 1, brute_force=True))


def test_ip_retrieval_function_matches_reference():
    for name in ('dense_ip_gauss', 'dense_ip_grid_k_gt_n'):
        g = load_golden(name)
        k = int(g['topk'])
        qids = ['q%d' % i for i in range(g['q_vals'].shape[0])]
        res, sc = gr.IP_retrieval(qids, torch.from_numpy(_q(g)), torch.from_numpy(g['c_vals'].astype(np.float32)), go.make_args(topk=k))
        for i, qid in enumerate(qids):
            assert len(res[qid]) == g['ref_rows'].shape[1]        # argsort[:k] never fails for k > N
            assert np.abs(np.array(sc[qid]) - g['ref_scores'][i]).max() < 1e-3
            if 'grid' in name:   # exact arithmetic: identical score lists, rows agree inside tie groups (argsort tie order is arbitrary)
                assert np.array_equal(np.array(sc[qid]), g['ref_scores'][i])
                assert tie_groups_equal(np.array(res[qid]), g['ref_rows'][i], np.array(sc[qid]))
            else:
                assert np.mean(np.array(res[qid]) == g['ref_rows'][i]) > 0.98


@pytest.mark.parametrize('tag,kw', [('theta', dict(theta=0.8)), ('theta_rerank', dict(theta=0.8, rerank=True, agip_topk=400)),
                                     ('ip', dict(theta=0.8, IP=True)), ('ip_rerank', dict(theta=0.8, IP=True, rerank=True, agip_topk=400))])
def test_approximate_modes_match_reference(tag, kw):
    g = load_golden('delade_approx_grid')
    S, k = int(g['S']), int(g['topk'])
    qids = list(range(g['q_vals'].shape[0]))
    res, sc = gr.GIP_retrieval(qids, torch.from_numpy(_q(g)), torch.from_numpy(g['q_idx']),
                               torch.from_numpy(g['c_vals'].astype(np.float32)), torch.from_numpy(g['c_idx']),
                               go.make_args(emb_dim=S, topk=k, **kw))
    got = np.array([sc[i] for i in qids])
    rows = np.array([res[i] for i in qids])
    if 'rerank' not in tag:
        # first-stage only: identical score lists (grid inputs), rows agree inside tie groups
        assert np.array_equal(got, g['ref_scores_' + tag])
        for i in qids:
            assert tie_groups_equal(rows[i], g['ref_rows_' + tag][i], got[i])
    else:
        # the reference's candidate set is cut at an arbitrary point inside a tie group (torch.topk), ours at the
        # lowest rows; every returned score must be the exact GIP score of that row and lists must agree where the
        # candidate cut does not interfere (top of the list)
        ex = go.gip_scores_f64(_q(g), g['q_idx'], g['c_vals'], g['c_idx'], S, 1)
        for i in qids:
            assert np.array_equal(ex[i][rows[i]], got[i])
            assert np.array_equal(got[i][:10], g['ref_scores_' + tag][i][:10])


def test_cli_main_round_trip(tmp_path, monkeypatch):
    g = load_golden('main_trec_grid')
    docids = [str(x) for x in g['docids']]
    qids = [str(x) for x in g['qids']]
    qp, ip = tmp_path / 'q.pt', tmp_path / 'c.index.pt'
    with open(qp, 'wb') as f:
        pickle.dump([g['q_vals'], g['q_idx'], qids], f, protocol=4)
    with open(ip, 'wb') as f:
        pickle.dump([g['c_vals'], g['c_idx'], docids], f, protocol=4)
    monkeypatch.chdir(tmp_path)
    base = ['--query_emb_path', str(qp), '--index_path', str(ip), '--emb_dim', str(int(g['S'])), '--brute_force', '--topk',
            str(int(g['topk'])), '--lamda', str(float(g['lamda'])), '--run_name', 'golden']

    def check(fn, ref_text):
        ours = open(fn).read().splitlines()
        ref = str(ref_text).splitlines()
        assert len(ours) == len(ref)
        key = lambda l: tuple(l.split(' ')[i] for i in (0, 1, 3, 4, 5))
        assert [key(l) for l in ours] == [key(l) for l in ref]    # qid, Q0, rank, score text, run name identical
        last = {l.split(' ')[0]: l.split(' ')[4] for l in ref}
        grp = lambda ls: {(l.split(' ')[0], l.split(' ')[4]): set() for l in ls}
        ga, gb = grp(ref), grp(ours)
        for l in ref:
            ga[(l.split(' ')[0], l.split(' ')[4])].add(l.split(' ')[2])
        for l in ours:
            gb[(l.split(' ')[0], l.split(' ')[4])].add(l.split(' ')[2])
        assert all(ga[kk] == gb[kk] for kk in ga if last[kk[0]] != kk[1])

    gr.main(base)
    check('result.trec', g['trec_single'])
    # same through the mmap-able .npy container (index_path is a directory)
    from dhr_b200 import index_io
    index_io.pickle_to_npy(str(ip), str(tmp_path / 'npy'))
    for sh in range(3):
        gr.main([a if a != str(ip) else str(tmp_path / 'npy') for a in base] + ['--total_shrad', '3', '--shrad', str(sh)])
        check('result%d.trec' % sh, g['trec_shard%d' % sh])
    for sh in range(3):
        gr.main(base + ['--total_shrad', '3', '--shrad', str(sh)])
        check('result%d.trec' % sh, g['trec_shard%d' % sh])
    # dense-only index through the CLI (idx entries None / 0)
    d = load_golden('main_trec_dense_grid')
    with open(qp, 'wb') as f:
        pickle.dump([d['q_vals'], None, [str(x) for x in d['qids']]], f, protocol=4)
    with open(ip, 'wb') as f:
        pickle.dump([d['c_vals'], 0, [str(x) for x in d['docids']]], f, protocol=4)
    gr.main(['--query_emb_path', str(qp), '--index_path', str(ip), '--topk', str(int(d['topk'])), '--run_name', 'golden'])
    check('result.trec', d['trec_single'])


def test_gpu_densify_op_matches_reference():
    from dhr_b200.densify import densify
    g = load_golden('densify_op')
    x = torch.from_numpy(g['x']).cuda()
    v, i = densify(x, dims=int(g['dims']), remove_dims=int(g['remove_dims']))
    assert np.array_equal(v.cpu().numpy(), g['ref_vals']) and np.array_equal(i.cpu().numpy(), g['ref_idx'])
    # reference-sized vocabulary, fp16 input, written into strided views of an index block
    rng = np.random.default_rng(3)
    big = np.maximum(rng.standard_normal((33, 30522)).astype(np.float32), 0).astype(np.float16)
    vals = torch.zeros((33, 896), dtype=torch.float16, device='cuda')
    idx = torch.zeros((33, 768), dtype=torch.uint8, device='cuda')
    densify(torch.from_numpy(big).cuda(), out=(vals[:, :768], idx))
    ev, ei = go.densify_port(big.astype(np.float32))
    assert np.array_equal(vals[:, :768].cpu().numpy(), ev) and np.array_equal(idx.cpu().numpy(), ei)
    assert torch.all(vals[:, 768:] == 0)
    with pytest.raises(ValueError):
        densify(x, dims=7)


@pytest.mark.parametrize('tag,rerank', [('rerank', True), ('norerank', False)])
def test_pq_ip_retrieval_rerank_half_matches_reference(tag, rerank, tmp_path):
    """a11: the reference's PQ_IP_retrieval (:167-231) ran on stub first-stage lists (tests/golden/make_golden.py pq);
    same lists -> same scores (grid inputs: bit-exact), rows equal inside tie groups."""
    g = load_golden('pq_rerank_grid')
    S, k = int(g['S']), int(g['topk'])
    qids = list(range(g['q_vals'].shape[0]))
    args = go.make_args(emb_dim=S, topk=k, agip_topk=int(g['agip_topk']), rerank=rerank, batch=4)
    served = []

    def first_stage(x, kk):                                           # faiss IndexPQ.search signature, batches in order
        lo = sum(served)
        served.append(x.shape[0])
        return g['candidate_scores'][lo:lo + x.shape[0], :kk], g['candidates'][lo:lo + x.shape[0], :kk]

    np.savez(tmp_path / 'cand.npz', candidates=g['candidates'], scores=g['candidate_scores'])
    file_args = go.make_args(emb_dim=S, topk=k, agip_topk=int(g['agip_topk']), rerank=rerank, faiss_pq_index_path=str(tmp_path / 'cand.npz'))
    for kw, a in ((dict(candidates=g['candidates'], candidate_scores=g['candidate_scores']), args), (dict(first_stage=first_stage), args),
                  ({}, file_args)):
        res, sc = gr.PQ_IP_retrieval(qids, torch.from_numpy(_q(g)), torch.from_numpy(g['q_idx']),
                                     torch.from_numpy(g['c_vals'].astype(np.float32)), torch.from_numpy(g['c_idx']), a, **kw)
        got = np.array([sc[i] for i in qids])
        rows = np.array([res[i] for i in qids])
        assert np.array_equal(got, g['ref_scores_' + tag])
        for i in qids:
            assert tie_groups_equal(rows[i], g['ref_rows_' + tag][i], got[i])
    assert served == [4, 2]


def test_host_output_staging_grows_per_dimension():
    """(Q=8, k=1000) then (Q=4096, k=1) with numpy outputs must not reuse a counts buffer sized for 8 queries."""
    case = make_case(5, 5000, 4096, 16, 2, 16, 39, np.uint8, grid=True)
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=16, group=2) as ix:
        s, r, c = ix.search(case['q_vals'][:8], case['q_idx'][:8], 1000)
        assert_matches_oracle(dict(case, q_vals=case['q_vals'][:8], q_idx=case['q_idx'][:8]), s, r, c, 1000, exact=True)
        s, r, c = ix.search(case['q_vals'], case['q_idx'], 1)
        assert np.all(c == 1)
        sub = slice(4000, 4096)
        assert_matches_oracle(dict(case, q_vals=case['q_vals'][sub], q_idx=case['q_idx'][sub]), s[sub], r[sub], c[sub], 1, exact=True)
        s, r, c = ix.search(case['q_vals'][:8], case['q_idx'][:8], 1000)          # and back: still intact
        assert_matches_oracle(dict(case, q_vals=case['q_vals'][:8], q_idx=case['q_idx'][:8]), s, r, c, 1000, exact=True)


def test_rerank_with_duplicate_candidates_is_memory_safe():
    """duplicated candidate rows break the unique-key assumption of the select; k a power of two is the case that
    used to write past the compaction buffer.  The duplicate may be returned twice; nothing may be corrupted."""
    case = make_case(9, 2000, 4, 16, 1, 16, 39, np.uint8, grid=True)
    k = 64
    cand = np.tile(np.arange(100, dtype=np.int64), (4, 2))                         # every candidate twice
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=16, group=1) as ix:
        s, r, c = ix.rerank(case['q_vals'], case['q_idx'], cand, k)
        ex = go.gip_scores_f64(case['q_vals'], case['q_idx'], case['c_vals'], case['c_idx'], 16, 1)
        for i in range(4):
            assert np.all((r[i] >= 0) & (r[i] < 100))
            assert np.array_equal(ex[i][r[i]].astype(np.float32), s[i])
            assert np.all(np.diff(s[i]) <= 0)
        s2, r2, c2 = ix.search(case['q_vals'], case['q_idx'], k)                    # the index still answers correctly
        assert_matches_oracle(case, s2, r2, c2, k, exact=True)


def test_rowmajor_arrays_are_dropped_and_rebuilt_on_demand():
    """only the tiled copies stay resident; K1 / rerank / fp32 queries rebuild the row-major arrays."""
    case = make_case(11, 6000, 9, 32, 6, 64, 39, np.uint16)
    k = 50
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=32, group=6, keep_rowmajor=True) as keep:
        full = keep.device_bytes
        s0, r0, c0 = keep.search(case['q_vals'], case['q_idx'], k)
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=32, group=6) as ix:
        lean = ix.device_bytes
        assert lean < 0.62 * full                                                   # tiled copies only (+ small workspaces)
        s, r, c = ix.search(case['q_vals'], case['q_idx'], k)                       # tile path: no rebuild
        assert ix.stats()['scan_variant'] == 3 and ix.stats()['rowmajor_rebuilds'] == 0
        assert np.array_equal(s, s0) and np.array_equal(r, r0)
        ix.set_option('tile_mode', 0)                                               # row scan K1 needs them
        s1, r1, c1 = ix.search(case['q_vals'], case['q_idx'], k)
        assert ix.stats()['rowmajor_rebuilds'] == 1
        assert_matches_oracle(case, s1, r1, c1, k)
        cand = r0[:, ::-1].copy()
        s2, r2, _ = ix.rerank(case['q_vals'], case['q_idx'], cand, k)               # K4 on the rebuilt arrays
        assert np.array_equal(np.sort(r2, axis=1), np.sort(r0, axis=1)) and np.allclose(s2, s0, atol=1e-5)
        before = ix.device_bytes
        ix.set_option('rowmajor', 0)
        assert ix.device_bytes <= before - 0.9 * 6000 * ix.row_bytes                # the three row-major arrays are gone again
        ix.set_option('tile_mode', 1)
        s3, r3, _ = ix.search(case['q_vals'], case['q_idx'], k)
        assert np.array_equal(s3, s0) and np.array_equal(r3, r0)
