"""CPU-only tests: host-side logic of the product, and that the C-ABI library loads and exports every symbol
include/dhr_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import gip_oracle as go


def test_cabi_exports_every_declared_symbol():
    from dhr_b200 import _cabi as C
    hdr = open(os.path.join(ROOT, 'include', 'dhr_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(dhr_[a-z_0-9]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    lib = C.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), 'libdhr_b200.so does not export ' + name
    assert declared == set(C.EXPORTS)
    assert lib.dhr_version() >= 100
    assert lib.dhr_strerror(0) == b'ok' and b'fp16' in lib.dhr_strerror(C.ERR_LOSSY)


def test_constants_match_header():
    from dhr_b200 import _cabi as C
    hdr = open(os.path.join(ROOT, 'include', 'dhr_b200.h')).read()
    defs = dict(re.findall(r'#define\s+(DHR_[A-Z_0-9]+)\s+(\d+)u?\b', hdr))
    assert int(defs['DHR_MAX_K']) == C.MAX_K and int(defs['DHR_MAX_GROUP']) == C.MAX_GROUP
    for name, val in [('DHR_IDX_U8', C.IDX_U8), ('DHR_IDX_I8', C.IDX_I8), ('DHR_IDX_I16', C.IDX_I16), ('DHR_IDX_U16', C.IDX_U16),
                      ('DHR_IDX_I32', C.IDX_I32), ('DHR_IDX_I64', C.IDX_I64), ('DHR_VAL_F16', C.VAL_F16), ('DHR_VAL_F32', C.VAL_F32),
                      ('DHR_ERR_LOSSY', C.ERR_LOSSY), ('DHR_ERR_IDX_RANGE', C.ERR_IDX_RANGE), ('DHR_ERR_NO_DEVICE', C.ERR_NO_DEVICE)]:
        assert int(defs[name]) == val
    assert ctypes.sizeof(C.DhrStats) == 10 * 4 + 8 * 8
    assert int(defs['DHR_INDEX_KEEP_ROWMAJOR']) == C.INDEX_KEEP_ROWMAJOR and int(defs['DHR_INDEX_NARROW_CODES']) == C.INDEX_NARROW_CODES


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from dhr_b200 import GipIndex, _cabi as C
    with pytest.raises(C.DhrError) as e:
        GipIndex.from_arrays(np.zeros((4, 8), np.float16), None)
    assert e.value.status == C.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'dhr_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('# oracle', ''), f + ' mentions the oracle'


def test_shard_bounds_match_reference_rule():
    from dhr_b200 import shard_bounds
    for n in (0, 1, 7, 400, 8841823):
        for t in (1, 2, 3, 8, 16):
            covered = []
            for s in range(t):
                assert shard_bounds(n, t, s) == go.shard_bounds(n, t, s)
                covered.append(shard_bounds(n, t, s))
            assert covered[0][0] == 0 and covered[-1][1] == n
            assert all(covered[i][1] == covered[i + 1][0] for i in range(t - 1))
    assert shard_bounds(8841823, 8, 7) == (7736589, 8841823)      # last shard takes the remainder (1,105,234 rows)


def test_write_trec_matches_reference_format(tmp_path):
    from dhr_b200 import write_trec
    g = load_golden('main_trec_grid')
    docids = [str(x) for x in g['docids']]
    qids = [str(x) for x in g['qids']]
    S, k, lam = int(g['S']), int(g['topk']), float(g['lamda'])
    q = g['q_vals'].astype(np.float32)
    q[:, -int(g['C']):] *= np.float32(lam)
    rows, vals = go.search_f64(q, g['q_idx'], g['c_vals'], g['c_idx'], S, 1, k)
    res = {qid: rows[i].tolist() for i, qid in enumerate(qids)}
    sc = {qid: vals[i].astype(np.float32).tolist() for i, qid in enumerate(qids)}
    out = tmp_path / 'result.trec'
    write_trec(str(out), res, sc, docids, 'golden')
    ours = out.read_text().splitlines()
    ref = str(g['trec_single']).splitlines()
    assert len(ours) == len(ref)
    key = lambda l: (l.split(' ')[0], l.split(' ')[1], l.split(' ')[3], l.split(' ')[4], l.split(' ')[5])
    assert [key(l) for l in ours] == [key(l) for l in ref]        # qid, Q0, rank (not renumbered), score text, run name


def test_synth_is_shard_consistent_and_has_reference_conventions():
    from dhr_b200 import synth
    full_v, full_i = synth.corpus_numpy('delade_cls', 0, 70000)
    a_v, a_i = synth.corpus_numpy('delade_cls', 65000, 70000)
    assert np.array_equal(full_v[65000:], a_v) and np.array_equal(full_i[65000:], a_i)
    S, G = 128, 6
    lex = full_v[:, :S * G].reshape(-1, S, G)
    empty = (lex == 0).all(axis=2)
    assert 0.25 < empty.mean() < 0.35 and np.all(full_i[empty] == 0)       # empty slice = value 0 AND idx 0
    assert full_i.max() < 39 and full_v.dtype == np.float16 and full_i.dtype == np.uint16
    qv, qi = synth.queries_numpy('bm25', 16)
    nz = (qv.reshape(16, 256, 3) != 0).any(axis=2).sum(axis=1)
    assert np.all(nz <= 8)


def test_cli_parser_accepts_reference_flags_verbatim():
    from dhr_b200.gip_retrieval import build_parser
    a = build_parser().parse_args(['--query_emb_path', 'q', '--index_path', 'i', '--emb_dim', '768', '--theta', '0.3', '--rerank',
                                   '--use_gpu', '--combine_cls', '--topk', '1000', '--total_shrad', '2', '--shrad', '1',
                                   '--lamda', '0.5', '--agip_topk', '10000', '--IP', '--brute_force', '--batch', '4',
                                   '--run_name', 'x', '--faiss_pq_index_path', 'p'])
    assert a.total_shrad == 2 and a.shrad == 1 and a.lamda == 0.5 and a.rerank and a.IP and a.brute_force


def test_index_container_round_trip(tmp_path):
    """pickle triple -> mmap-able .npy directory -> pickle triple is lossless; shards map only their rows."""
    import pickle
    from dhr_b200 import index_io
    g = load_golden('main_trec_grid')
    docids = [str(x) for x in g['docids']]
    src = tmp_path / 'c.index.pt'
    with open(src, 'wb') as f:
        pickle.dump([g['c_vals'], g['c_idx'], docids], f, protocol=4)
    d = tmp_path / 'npy'
    index_io.main(['to-npy', str(src), str(d)])
    back = tmp_path / 'back.index.pt'
    index_io.main(['to-pickle', str(d), str(back)])
    with open(back, 'rb') as f:
        v, i, ids = pickle.load(f)
    assert np.array_equal(v, g['c_vals']) and v.dtype == np.float16
    assert np.array_equal(i, g['c_idx']) and i.dtype == g['c_idx'].dtype and ids == docids
    for sh in range(3):
        vals, idx, ids, lo = index_io.load_npy(str(d), 3, sh)
        a, b = go.shard_bounds(len(docids), 3, sh)
        assert lo == a and np.array_equal(vals, g['c_vals'][a:b]) and np.array_equal(idx, g['c_idx'][a:b]) and ids == docids[a:b]
    # dense-only index: the reference stores the int 0 in place of idx (index.py:40-43)
    with open(src, 'wb') as f:
        pickle.dump([g['c_vals'], 0, list(range(len(docids)))], f, protocol=4)
    index_io.main(['to-npy', str(src), str(tmp_path / 'dense')])
    vals, idx, ids, _ = index_io.load_npy(str(tmp_path / 'dense'))
    assert idx is None and ids == list(range(len(docids)))


def test_merge_splits_matches_reference_semantics(tmp_path):
    import pickle
    from dhr_b200 import index_io
    g = load_golden('main_trec_grid')
    n = g['c_vals'].shape[0]
    cuts = [0, 150, 290, n]
    for k in range(3):
        with open(tmp_path / ('corpus.split%02d.pt' % k), 'wb') as f:
            pickle.dump([g['c_vals'][cuts[k]:cuts[k + 1]], g['c_idx'][cuts[k]:cuts[k + 1]],
                         [str(x) for x in g['docids'][cuts[k]:cuts[k + 1]]]], f, protocol=4)
    out = index_io.merge_splits(str(tmp_path), 'corpus')
    with open(out, 'rb') as f:
        v, i, ids = pickle.load(f)
    assert np.array_equal(v, g['c_vals']) and np.array_equal(i, g['c_idx']) and ids == [str(x) for x in g['docids']]


def _py_write_trec(path, results, scores, docids, run_name):
    """the reference's formatting loop (gip_retrieval.py:329-342), kept here as the expected text"""
    with open(path, 'w') as fout:
        for query_id in results:
            for rank, docidx in enumerate(results[query_id]):
                if docids[docidx] != query_id:
                    fout.write('{} Q0 {} {} {} {}\n'.format(query_id, docids[docidx], rank + 1, scores[query_id][rank], run_name))


@pytest.mark.parametrize('ids', ['int', 'str', 'mixed'])
def test_trec_writer_matches_python_formatting(tmp_path, ids):
    """C++ TREC writer (csrc/trec.cu): byte-identical to the reference's Python loop, including float repr, the
    docid == qid skip rule without renumbering, and ragged per-query lengths."""
    from dhr_b200.gip_retrieval import write_trec
    rng = np.random.default_rng(5)
    n_docs, nq, k = 500, 37, 60
    special = np.array([0.0, -0.0, 1.0, 100.0, 1e-5, 9.999999e-5, 1e-4, 123456.789, 1e16, 9.9e15, 3.4028235e38, 1.17549435e-38,
                        1e-45, 0.1, 0.3, 16777216.0, 2.5e-7, -7.25, 65504.0, 1e7], np.float32)
    if ids == 'int':
        docids = list(range(1000, 1000 + n_docs)); qids = [int(x) for x in rng.choice(np.arange(900, 1600), nq, replace=False)]
    elif ids == 'str':
        docids = ['D%d' % i for i in range(n_docs)]; qids = ['D%d' % i for i in rng.choice(n_docs * 2, nq, replace=False)]
    else:
        docids = list(range(n_docs)); qids = [str(i) for i in range(nq)]        # int docids vs str qids never compare equal
    results, scores = {}, {}
    for q in qids:
        n = int(rng.integers(1, k + 1))
        rows = rng.choice(n_docs, n, replace=False)
        if ids != 'mixed' and q in docids and rng.random() < 0.8:
            rows[rng.integers(0, n)] = docids.index(q)                             # force the skip rule
        sc = np.sort(np.concatenate([rng.standard_normal(n).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 7)), special]))[::-1][:n]
        results[q] = rows.tolist()
        scores[q] = sc.astype(np.float32).tolist()
    a, b = str(tmp_path / 'a.trec'), str(tmp_path / 'b.trec')
    write_trec(a, results, scores, docids, 'h2oloo')
    _py_write_trec(b, results, scores, docids, 'h2oloo')
    ta, tb = open(a).read(), open(b).read()
    assert ta == tb
    assert len(ta) > 0


def test_merge_result_cli(tmp_path, monkeypatch):
    """3 shard files written by the CLI merge into the single-shard result (tie groups as sets)."""
    from dhr_b200 import merge_result
    g = load_golden('main_trec_grid')
    monkeypatch.chdir(tmp_path)
    for sh in range(3):
        with open('result%d.trec' % sh, 'w') as f:
            f.write(str(g['trec_shard%d' % sh]))
    merge_result.main(['--total_shrad', '3', '--topk', str(int(g['topk'])), '--run_name', 'golden'])
    ours = open('result.trec').read().splitlines()
    # expected: per query, the best topk of the union by score; compare (qid, rank, score) sequences with an oracle merge
    exp = {}
    for sh in range(3):
        for l in str(g['trec_shard%d' % sh]).splitlines():
            f = l.split(' ')
            exp.setdefault(f[0], []).append((float(f[4]), f[2]))
    k = int(g['topk'])
    got = {}
    for l in ours:
        f = l.split(' ')
        got.setdefault(f[0], []).append((float(f[4]), f[2], int(f[3])))
    for q, items in exp.items():
        want = sorted((s for s, _ in items), reverse=True)[:k]
        assert [s for s, _, _ in got[q]] == want
        assert [r for _, _, r in got[q]] == list(range(1, len(want) + 1))
        assert all((s, d) in items for s, d, _ in got[q])


def test_merge_trec_matches_python_rule(tmp_path):
    """C++ shard merge (csrc/trec.cu): same text as a Python statement of the rule -- per query the union of the shards'
    lines, (score desc, position asc), ranks from 1, scores printed as Python prints float(text) -- including ties, ragged
    shards, queries missing from a shard and exponent-form scores."""
    from dhr_b200.merge_result import merge_trec
    rng = np.random.default_rng(11)
    qids = ['q%d' % i for i in range(23)]
    paths = []
    per_q = {}
    for sh in range(4):
        p = str(tmp_path / ('result%d.trec' % sh))
        paths.append(p)
        with open(p, 'w') as f:
            for q in qids:
                if rng.random() < 0.15:
                    continue                                             # this shard has nothing for the query
                n = int(rng.integers(1, 40))
                sc = np.sort(np.round(rng.standard_normal(n).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 5)), 3))[::-1]
                sc[rng.random(n) < 0.3] = np.float32(1.5)                # ties across and inside shards
                sc = np.sort(sc)[::-1]
                for r in range(n):
                    doc = 'D%d_%d' % (sh, int(rng.integers(0, 10 ** 6)))
                    text = repr(float(sc[r]))
                    f.write('%s Q0 %s %d %s shard\n' % (q, doc, r + 1, text))
                    per_q.setdefault(q, []).append((float(text), doc))
    out = str(tmp_path / 'merged.trec')
    k = 25
    n_lines = merge_trec(paths, out, k, 'dhr')
    exp = []
    for q, items in per_q.items():                                       # dict order = order of first appearance
        order = sorted(range(len(items)), key=lambda i: (-items[i][0], i))[:k]
        for rank, i in enumerate(order):
            exp.append('{} Q0 {} {} {} {}\n'.format(q, items[i][1], rank + 1, items[i][0], 'dhr'))
    got = open(out).read()
    assert got == ''.join(exp)
    assert n_lines == len(exp)


def test_k1t_postings_spec_matches_oracle():
    """tools/k1t_postings_spec.py (the layout + walk the next K1t is specified by): code-sorted postings inside a tile,
    query-driven walk == the exact lexical scores, on exact-arithmetic inputs incl. empty slices and a ragged last tile."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('k1t_postings_spec', os.path.join(ROOT, 'tools', 'k1t_postings_spec.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from helpers import make_case
    from oracle import c_oracle
    S, G, R = 16, 3, 12
    case = make_case(3, 700, 9, S, G, 0, R, np.uint8, np.uint8, c_density=0.6, q_density=0.7, grid=True)
    tiles = mod.build_postings(case['c_vals'], case['c_idx'], S, G, R, tile_rows=512)
    assert [t['rows'] for t in tiles] == [512, 188]
    got = np.concatenate([mod.walk_tile(t, case['q_vals'], case['q_idx'], S, G, R)[0] for t in tiles], axis=1)
    for q in range(case['q_vals'].shape[0]):
        ex = c_oracle.scores(case['c_vals'], case['c_idx'], case['q_vals'][q].astype(np.float32), case['q_idx'][q], S, G)
        assert np.array_equal(got[q].astype(np.float64), ex)
    # every stored item is a non-empty slice and lists are sorted by (code, passage)
    t0 = tiles[0]
    assert np.all(np.any(t0['val'] != 0, axis=1))
    for s in range(S):
        for c in range(R):
            p = t0['pid'][t0['off'][s, c]:t0['off'][s, c + 1]]
            assert np.all(np.diff(p.astype(np.int64)) > 0)
            assert np.all(case['c_idx'][p, s] == c)


def test_packed_keys_roundtrip_and_order():
    """exchange format of the sharded search: descending int64-as-uint64 key order == (score desc, row asc)"""
    from dhr_b200 import pack_keys, unpack_keys
    rng = np.random.default_rng(0)
    s = np.concatenate([rng.standard_normal(500).astype(np.float32), np.float32([0.0, -0.0, 1.5, 1.5, 1.5, -3.25, -3.25])])
    r = rng.permutation(len(s)).astype(np.int64) + 8_000_000
    k = pack_keys(s, r)
    s2, r2 = unpack_keys(k)
    assert np.array_equal(r2, r) and np.array_equal(s2, s + np.float32(0.0))
    order = np.argsort(k.view(np.uint64))[::-1]
    expect = np.lexsort((r, -(s.astype(np.float64))))
    assert np.array_equal(order, expect)
    pad = pack_keys(np.float32([1.0]), np.int64([-1]))
    assert pad[0] == 0 and unpack_keys(pad)[1][0] == -1 and np.isneginf(unpack_keys(pad)[0][0])


def test_trec_writer_longest_admitted_ids_and_limits(tmp_path):
    """a 500-character query id, a 399-character docid and a 250-character run name used to overrun the line
    buffer of dhr_write_trec.  The admitted maxima (512 / 400 / 256) format correctly; anything longer is rejected."""
    from dhr_b200.gip_retrieval import write_trec_arrays
    from dhr_b200 import _cabi as C
    qids = ['q' * 500, 'short']
    docids = ['d' * 399, 'e' * 400, 'x']
    rows = np.array([[0, 1, 2], [2, 1, -1]], np.int64)
    scores = np.array([[3.5, 2.25, -1.0], [0.1, 1e-5, 0.0]], np.float32)
    out = str(tmp_path / 'long.trec')
    run = 'r' * 256
    n = write_trec_arrays(out, qids, scores, rows, None, docids, run)
    assert n == 5
    lines = open(out).read().splitlines()
    exp = ['{} Q0 {} {} {} {}'.format(qids[q], docids[rows[q, r]], r + 1, float(scores[q, r]), run)
           for q in range(2) for r in range(3) if rows[q, r] >= 0]
    assert lines == exp
    with pytest.raises(C.DhrError):
        write_trec_arrays(out, qids, scores, rows, None, docids, 'r' * 257)
    with pytest.raises(C.DhrError):
        write_trec_arrays(out, ['q' * 513, 'short'], scores, rows, None, docids, 'run')
    with pytest.raises(C.DhrError):
        write_trec_arrays(out, qids, scores, rows, None, ['d' * 401, 'e', 'x'], 'run')
