"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors produced by the real
reference and against the CPU oracle on seeded inputs.  Run on the B200 box: pytest -m gpu."""
import ctypes

import numpy as np
import pytest

from conftest import load_golden, has_cuda
from helpers import make_case, assert_matches_oracle, tie_groups_equal

pytestmark = pytest.mark.gpu

if has_cuda():
    import torch
    from dhr_b200 import GipIndex, topk_merge, _cabi as C
    from oracle import gip_oracle as go


def configure(ix, mode, qb=None):
    """mode: 'scan0' / 'scan1' = row-scan kernel K1 (direct loads / TMA bulk staging), 'tile' = tensor-core +
    bucketed tile kernels (K2 + K1t) when the shape allows (falls back to K1 otherwise)."""
    if mode in ('tile', 'tile_p'):
        ix.set_option('tile_mode', 1)
    else:
        ix.set_option('tile_mode', 0)
        ix.set_option('scan_variant', int(mode[-1]))
    if qb:
        ix.set_option('query_block', qb)


MODES = ['scan0', 'scan1', 'tile', 'tile_p']     # tile = tiled lexical layout (K1t), tile_p = experimental postings layout (K1p)


def postings(variant):
    return variant == 'tile_p'


def _golden_queries(g):
    return g['q_vals'].astype(np.float32)


GRID = ['delade_g1_u8_grid', 'bm25_i16_grid', 'unicoil_i8_i16_grid', 'grouped_g6_u16_grid', 'grouped_g3_u16_grid', 'bm25_wide_i16_grid',
        'grouped_g3_wide_u16_grid',
        'delade_lamda_grid']


@pytest.mark.parametrize('variant', MODES)
@pytest.mark.parametrize('name', GRID)
def test_golden_grid_bit_exact(name, variant):
    """Grid fixtures: the reference's fp32 scores are order-independent, so our score lists must be
    bit-identical to the reference's and rows must agree inside tie groups."""
    g = load_golden(name)
    S, G, k = int(g['S']), int(g['G']), int(g['topk'])
    lam = float(g['lamda']) if 'lamda' in g else 1.0
    with GipIndex.from_arrays(g['c_vals'], g['c_idx'], n_slices=S, group=G, lex_postings=postings(variant)) as ix:
        configure(ix, variant)
        scores, rows, counts = ix.search(_golden_queries(g), g['q_idx'], k, lamda=lam)
    assert np.array_equal(scores.astype(np.float64), g['ref_scores'])
    for i in range(rows.shape[0]):
        assert tie_groups_equal(rows[i], g['ref_rows'][i], scores[i])
    case = dict(S=S, G=G, C=int(g['C']), c_vals=g['c_vals'], c_idx=g['c_idx'], q_vals=g['q_vals'], q_idx=g['q_idx'])
    assert_matches_oracle(case, scores, rows, counts, k, lamda=lam, exact=True)


@pytest.mark.parametrize('variant', MODES)
def test_golden_gauss_within_tolerance(variant):
    g = load_golden('delade_g1_u8_gauss')
    S, G, k = int(g['S']), int(g['G']), int(g['topk'])
    with GipIndex.from_arrays(g['c_vals'], g['c_idx'], n_slices=S, group=G, lex_postings=postings(variant)) as ix:
        configure(ix, variant)
        scores, rows, counts = ix.search(_golden_queries(g), g['q_idx'], k)
    assert np.abs(scores - g['ref_scores']).max() < 1e-3          # north_star tolerance
    assert np.mean(rows == g['ref_rows']) > 0.99                  # ranks identical outside fp32-noise ties
    case = dict(S=S, G=G, C=int(g['C']), c_vals=g['c_vals'], c_idx=g['c_idx'], q_vals=g['q_vals'], q_idx=g['q_idx'])
    assert_matches_oracle(case, scores, rows, counts, k)


@pytest.mark.parametrize('name', ['dense_ip_gauss', 'dense_ip_grid_k_gt_n'])
def test_golden_dense_only(name):
    g = load_golden(name)
    k = int(g['topk'])
    n = g['c_vals'].shape[0]
    with GipIndex.from_arrays(g['c_vals'], None) as ix:
        scores, rows, counts = ix.search(_golden_queries(g), None, min(k, C.MAX_K))
    kk = min(k, n)
    assert np.all(counts == kk)
    assert np.abs(scores[:, :kk] - g['ref_scores']).max() < 1e-3
    case = dict(S=0, G=1, C=int(g['C']), c_vals=g['c_vals'], c_idx=None, q_vals=g['q_vals'], q_idx=None)
    assert_matches_oracle(case, scores, rows, counts, min(k, C.MAX_K))


SHAPES = [
    # S, G, C, R, c_idx dtype, q_idx dtype
    (64, 1, 32, 39, np.uint8, np.uint8),
    (40, 1, 0, 39, np.int8, np.int16),          # S not a multiple of 16 -> padded slices
    (48, 2, 16, 100, np.int16, np.int16),
    (32, 3, 0, 3466, np.uint16, np.uint16),
    (24, 4, 24, 39, np.uint8, np.int64),
    (16, 5, 8, 39, np.int32, np.int32),
    (128, 6, 768, 39, np.uint16, np.uint16),    # BASELINE config 2 row shape
    (16, 7, 100, 39, np.uint8, np.uint8),       # C not a multiple of 8 -> padded columns
    (24, 8, 40, 39, np.uint16, np.uint8),
    (0, 1, 768, 1, np.uint8, np.uint8),         # dense only
    (768, 1, 128, 39, np.uint8, np.uint8),      # reference-true DeLADE shape
]


@pytest.mark.parametrize('variant', MODES)
@pytest.mark.parametrize('qb', [1, 2, 4, 8])
@pytest.mark.parametrize('shape', SHAPES)
def test_oracle_parity_shapes(shape, qb, variant):
    S, G, Cd, R, cdt, qdt = shape
    if variant in ('tile', 'tile_p') and qb != 1:
        pytest.skip('query_block only applies to the row scan')
    case = make_case(100 + S + G + Cd, 3000, 11, S, G, Cd, R, cdt, qdt)
    k = 100
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'] if S else None, n_slices=S, group=G, lex_postings=postings(variant)) as ix:
        configure(ix, variant, qb)
        scores, rows, counts = ix.search(case['q_vals'], case['q_idx'] if S else None, k)
    assert_matches_oracle(case, scores, rows, counts, k)


@pytest.mark.parametrize('variant', MODES)
@pytest.mark.parametrize('qb', [1, 4])
def test_fp32_queries_lamda_and_unmasked(qb, variant):
    case = make_case(7, 4000, 9, 64, 3, 40, 39, np.uint8, np.int16, q_fp32_noise=True)
    k = 50
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=64, group=3, lex_postings=postings(variant)) as ix:
        configure(ix, variant, qb)
        s1, r1, c1 = ix.search(case['q_vals'], case['q_idx'], k, lamda=0.37)
        s2, r2, c2 = ix.search(case['q_vals'], None, k, masked=False)
    assert_matches_oracle(case, s1, r1, c1, k, lamda=0.37)
    assert_matches_oracle(case, s2, r2, c2, k, masked=False)


def test_grid_many_shapes_bit_exact():
    """Exact-arithmetic inputs: every kernel variant must give the identical, deterministic answer."""
    for (S, G, Cd) in [(32, 1, 16), (16, 6, 64), (32, 3, 0)]:
        case = make_case(5 + G, 5000, 10, S, G, Cd, 8, np.uint8, np.uint8, c_density=0.5, q_density=0.5, grid=True)
        k = 200
        outs = []
        with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=S, group=G) as ix:
            for variant in ('scan0', 'scan1'):
                for qb in (1, 2, 4, 8):
                    configure(ix, variant, qb)
                    outs.append(ix.search(case['q_vals'], case['q_idx'], k))
            configure(ix, 'tile')
            outs.append(ix.search(case['q_vals'], case['q_idx'], k))
            assert ix.stats()['scan_variant'] == 3 and ix.stats()['lex_layout'] == 0, 'tiled tile path not taken'
        with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=S, group=G, lex_postings=True) as ix:
            outs.append(ix.search(case['q_vals'], case['q_idx'], k))
            assert ix.stats()['scan_variant'] == 3 and ix.stats()['lex_layout'] == 1, 'postings tile path not taken'
        assert_matches_oracle(case, *outs[0], k, exact=True)
        for o in outs[1:]:
            assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])


def test_edge_cases_k_gt_n_single_row_all_ties():
    case = make_case(3, 37, 5, 16, 1, 8, 5, grid=True)
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=16) as ix:
        s, r, c = ix.search(case['q_vals'], case['q_idx'], 100)
        assert_matches_oracle(case, s, r, c, 100, exact=True)
    one = make_case(4, 1, 3, 16, 1, 8, 5, grid=True)
    with GipIndex.from_arrays(one['c_vals'], one['c_idx'], n_slices=16) as ix:
        s, r, c = ix.search(one['q_vals'], one['q_idx'], 1)
        assert_matches_oracle(one, s, r, c, 1, exact=True)
    # all scores equal (zero queries): ties must resolve to the lowest rows
    z = make_case(5, 50000, 3, 16, 1, 8, 5, grid=True)
    z['q_vals'][:] = 0
    with GipIndex.from_arrays(z['c_vals'], z['c_idx'], n_slices=16) as ix:
        s, r, c = ix.search(z['q_vals'], z['q_idx'], 1000)
    assert np.all(s == 0) and np.array_equal(r, np.tile(np.arange(1000), (3, 1)))


def test_adversarial_ascending_scores_use_overflow_fallback():
    """Scores increase with the row index, so every row beats the running threshold: the candidate buffer
    overflows and the query is re-run with the overflow-proof schedule.  Result must still be exact."""
    n = 120000
    c_vals = np.zeros((n, 8), np.float16)
    c_vals[:, 0] = (np.arange(n) // 64).astype(np.float16)        # non-decreasing, many ties, exact in fp16
    q_vals = np.zeros((2, 8), np.float32)
    q_vals[0, 0] = 1.0
    q_vals[1, 0] = -1.0                                           # descending for the second query: no overflow
    case = dict(S=0, G=1, C=8, c_vals=c_vals, c_idx=None, q_vals=q_vals, q_idx=None)
    with GipIndex.from_arrays(c_vals, None) as ix:
        s, r, c = ix.search(q_vals, None, 1000)
        st = ix.stats()
    assert st['n_fallback_queries'] >= 1
    assert_matches_oracle(case, s, r, c, 1000, exact=True)


@pytest.mark.parametrize('k', [1, 1000, 4096, 10000])
def test_k_range(k):
    case = make_case(11, 60000, 3, 32, 1, 16, 39)
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=32) as ix:
        s, r, c = ix.search(case['q_vals'], case['q_idx'], k)
    assert_matches_oracle(case, s, r, c, k)


def test_device_pointers_and_strided_inputs():
    case = make_case(13, 5000, 6, 32, 2, 16, 39, np.int16, np.int16)
    dev = torch.device('cuda', 0)
    cv = torch.from_numpy(case['c_vals']).to(dev)
    ci = torch.from_numpy(case['c_idx']).to(dev)
    wide = torch.zeros((6, case['q_vals'].shape[1] + 5), dtype=torch.float32, device=dev)
    wide[:, :case['q_vals'].shape[1]] = torch.from_numpy(case['q_vals']).to(dev)
    qv = wide[:, :case['q_vals'].shape[1]]                        # row stride > width
    qi = torch.from_numpy(case['q_idx']).to(dev)
    with GipIndex.from_arrays(cv, ci, n_slices=32, group=2) as ix:
        s, r, c = ix.search(qv, qi, 64)
        assert s.is_cuda and r.is_cuda
        assert_matches_oracle(case, s.cpu().numpy(), r.cpu().numpy(), c.cpu().numpy(), 64)
        # fp16 queries from the host give the same answer
        s2, r2, c2 = ix.search(case['q_vals'].astype(np.float16), case['q_idx'], 64)
    assert np.array_equal(r2, r.cpu().numpy())


def test_fp32_corpus_lossless_and_lossy():
    case = make_case(17, 2000, 4, 16, 1, 8, 39)
    with GipIndex.from_arrays(case['c_vals'].astype(np.float32), case['c_idx'], n_slices=16) as ix:
        s, r, c = ix.search(case['q_vals'], case['q_idx'], 20)
    assert_matches_oracle(case, s, r, c, 20)
    bad = case['c_vals'].astype(np.float32)
    bad[5, 3] = 0.1                                               # not an fp16 number
    with pytest.raises(C.DhrError) as e:
        GipIndex.from_arrays(bad, case['c_idx'], n_slices=16)
    assert e.value.status == C.ERR_LOSSY


def test_error_paths():
    case = make_case(19, 100, 2, 16, 1, 8, 39)
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=16) as ix:
        with pytest.raises(C.DhrError) as e:
            ix.search(case['q_vals'], case['q_idx'], C.MAX_K + 1)
        assert e.value.status == C.ERR_UNSUPPORTED
        with pytest.raises(ValueError):
            ix.search(case['q_vals'][:, :-1], case['q_idx'], 5)
        with pytest.raises(C.DhrError):
            ix.search(case['q_vals'], case['q_idx'], 0)
    big = case['c_idx'].astype(np.uint8).copy()
    big[0, 0] = 255
    cv = case['c_vals'].copy()
    cv[0, 0] = 1.0
    with pytest.raises(C.DhrError) as e:
        GipIndex.from_arrays(cv, big, n_slices=16)
    assert e.value.status == C.ERR_IDX_RANGE
    ix = GipIndex(16, 8, capacity=10)
    with pytest.raises(C.DhrError) as e:                          # search before finalize
        ix.search(case['q_vals'], case['q_idx'], 5)
    assert e.value.status == C.ERR_STATE
    ix.close()


def test_rerank_and_merge():
    case = make_case(23, 8000, 5, 32, 1, 16, 39, grid=True)
    k = 40
    rng = np.random.default_rng(0)
    cand = np.stack([rng.choice(8000, size=500, replace=False) for _ in range(5)]).astype(np.int64)
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=32) as ix:
        s, r, c = ix.rerank(case['q_vals'], case['q_idx'], cand, k)
        full = ix.search(case['q_vals'], case['q_idx'], k)
    ex = go.gip_scores_f64(case['q_vals'], case['q_idx'], case['c_vals'], case['c_idx'], 32, 1)
    for i in range(5):
        sub = ex[i][cand[i]]
        order = np.lexsort((cand[i], -sub))[:k]
        assert np.array_equal(r[i], cand[i][order]) and np.array_equal(s[i].astype(np.float64), sub[order])
    # shard merge == single shard (same tie rule)
    parts_s, parts_r = [], []
    for sh in range(3):
        lo, hi = go.shard_bounds(8000, 3, sh)
        with GipIndex.from_arrays(case['c_vals'][lo:hi], case['c_idx'][lo:hi], n_slices=32, row_offset=lo) as ix:
            ps, pr, _ = ix.search(case['q_vals'], case['q_idx'], k)
        parts_s.append(ps)
        parts_r.append(pr)
    ms, mr = topk_merge(np.stack(parts_s), np.stack(parts_r))
    assert np.array_equal(ms, full[0]) and np.array_equal(mr, full[1])
    ms2, mr2 = topk_merge(torch.from_numpy(np.stack(parts_s)).cuda(), torch.from_numpy(np.stack(parts_r)).cuda())
    assert np.array_equal(ms2.cpu().numpy(), full[0]) and np.array_equal(mr2.cpu().numpy(), full[1])


def test_medium_config2_shape_against_c_oracle():
    """BASELINE config 2 row shape at 200k rows, 12 queries, top-1000, every query block and both variants."""
    from dhr_b200 import synth
    cv, ci = synth.corpus_numpy('delade_cls', 0, 200000)
    qv, qi = synth.queries_numpy('delade_cls', 12)
    case = dict(S=128, G=6, C=768, c_vals=cv, c_idx=ci, q_vals=qv.astype(np.float32), q_idx=qi)
    with GipIndex.from_arrays(cv, ci, n_slices=128, group=6) as ix:
        assert ix.row_bytes == 3328
        base = None
        for variant in ('scan0', 'scan1'):
            for qb in (1, 4, 8):
                configure(ix, variant, qb)
                out = ix.search(qv, qi, 1000)
                if base is None:
                    base = out
                    assert_matches_oracle(case, *out, 1000)
                else:
                    assert np.array_equal(out[1], base[1]) and np.array_equal(out[0], base[0])
        # tile kernels: different (deterministic) summation order -> compare through the oracle
        configure(ix, 'tile')
        out = ix.search(qv, qi, 1000)
        assert ix.stats()['scan_variant'] == 3, 'tile path not taken'
        assert_matches_oracle(case, *out, 1000)
        assert np.abs(out[0] - base[0]).max() < 1e-4


@pytest.mark.parametrize('cdim', [768, 64, 100, 200])
def test_dense_tile_tensor_core_path(cdim):
    """Dense-only index: the tcgen05 tile kernel (K2) must agree with the oracle and with the SIMT row scan."""
    case = make_case(31 + cdim, 70000, 150, 0, 1, cdim, 1)
    k = 100
    with GipIndex.from_arrays(case['c_vals'], None) as ix:
        ix.set_option('tile_mode', 1)
        s1, r1, c1 = ix.search(case['q_vals'], None, k)
        st = ix.stats()
        ix.set_option('tile_mode', 0)
        s0, r0, c0 = ix.search(case['q_vals'], None, k)
    assert st['scan_variant'] == 2, 'tile path not taken'
    assert_matches_oracle(case, s1, r1, c1, k)
    assert np.abs(s1 - s0).max() < 1e-4
    assert np.mean(r1 == r0) > 0.995


def test_dense_tile_grid_bit_exact_and_ties():
    case = make_case(37, 40000, 70, 0, 1, 128, 1, grid=True)
    k = 500
    with GipIndex.from_arrays(case['c_vals'], None) as ix:
        s1, r1, c1 = ix.search(case['q_vals'], None, k)
        assert ix.stats()['scan_variant'] == 2
    assert_matches_oracle(case, s1, r1, c1, k, exact=True)


@pytest.mark.parametrize('shape', [(128, 6, 768, 39, np.uint16), (768, 1, 128, 39, np.uint8), (64, 3, 0, 200, np.uint16),
                                   (40, 1, 32, 39, np.int8), (24, 8, 40, 39, np.uint16), (16, 5, 8, 39, np.int32),
                                   (16, 7, 100, 39, np.uint8), (48, 2, 16, 100, np.int16), (24, 4, 24, 39, np.uint8),
                                   # index range > 254: 16-bit codes in the tiled copy, distinct-code lookup ("wide" layout)
                                   (64, 3, 0, 3466, np.uint16), (32, 6, 64, 300, np.int16), (128, 1, 0, 1000, np.int16),
                                   (20, 2, 24, 40000, np.int32), (16, 5, 16, 500, np.uint16)])
@pytest.mark.parametrize('layout', ['postings', 'tiled'])
def test_tile_path_many_queries(shape, layout):
    """Tile kernels with several query tiles in flight (300 queries -> 5 tiles of 64, two super-batches of 256)."""
    S, G, Cd, R, cdt = shape
    case = make_case(900 + S + G, 60000, 300, S, G, Cd, R, cdt, cdt)
    k = 100
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=S, group=G, lex_postings=layout == 'postings') as ix:
        configure(ix, 'tile')
        s, r, c = ix.search(case['q_vals'], case['q_idx'], k)
        # shapes whose stage does not fit the K1t shared-memory budget (large G) fall back to the row scan
        if G <= 6:
            assert ix.stats()['scan_variant'] == 3, 'tile path not taken'
    sub = dict(case)
    sel = np.r_[0:8, 120:136, 250:260, 292:300]
    sub['q_vals'], sub['q_idx'] = case['q_vals'][sel], case['q_idx'][sel]
    assert_matches_oracle(sub, s[sel], r[sel], c[sel], k)


@pytest.mark.parametrize('shape', [(256, 3, 3466), (768, 1, 3466)])
def test_tile_path_bm25_like(shape):
    """Densified-BM25 shape (BASELINE config 3): 5 % of the corpus slices set, <= 8 query slices set with small-integer
    term frequencies (exact ties), idx up to 3465 -> wide tile layout; exact arithmetic -> bit-exact against the oracle."""
    S, G, R = shape
    rng = np.random.default_rng(77 + S)
    n, nq, k = 40000, 70, 100
    cv = np.zeros((n, S, G), np.float32)
    ci = np.zeros((n, S), np.int64)
    on = rng.random((n, S)) < 0.05
    cv[on] = rng.integers(32, 512, size=(int(on.sum()), G)).astype(np.float32) / 64.0
    ci[on] = rng.integers(0, R, size=int(on.sum()))
    qv = np.zeros((nq, S, G), np.float32)
    qi = np.zeros((nq, S), np.int64)
    for q in range(nq):
        sl = rng.choice(S, size=rng.integers(1, 9), replace=False)
        qv[q, sl] = rng.integers(1, 4, size=(len(sl), G)).astype(np.float32)
        # reuse indices that occur in the corpus so that matches exist
        qi[q, sl] = ci[rng.integers(0, n, size=len(sl)), sl]
    case = dict(S=S, G=G, C=0, c_vals=cv.reshape(n, S * G).astype(np.float16), c_idx=ci.astype(np.uint16),
                q_vals=qv.reshape(nq, S * G), q_idx=qi.astype(np.int16))
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=S, group=G) as ix:
        configure(ix, 'tile')
        s, r, c = ix.search(case['q_vals'], case['q_idx'], k)
        assert ix.stats()['scan_variant'] == 3, 'tile path not taken'
        configure(ix, 'scan1', 8)
        s2, r2, c2 = ix.search(case['q_vals'], case['q_idx'], k)
    assert_matches_oracle(case, s, r, c, k, exact=True)
    assert np.array_equal(s, s2) and np.array_equal(r, r2)


def test_dense_only_cta_pair_variant():
    """K2 as a CTA pair (cta_group::2): same answer as the single-CTA kernel on a dense-only index, > 128 queries in flight."""
    case = make_case(77, 70000, 300, 0, 1, 768, 1, np.uint8, grid=True)
    k = 100
    with GipIndex.from_arrays(case['c_vals'], None) as ix:
        s1, r1, c1 = ix.search(case['q_vals'], None, k)
        ix.set_option('dense_variant', 2)
        s2, r2, c2 = ix.search(case['q_vals'], None, k)
        assert ix.stats()['scan_variant'] == 2
    assert np.array_equal(s1, s2) and np.array_equal(r1, r2)
    sub = dict(case)
    sel = np.r_[0:4, 126:132, 296:300]
    sub['q_vals'], sub['q_idx'] = case['q_vals'][sel], None
    assert_matches_oracle(sub, s2[sel], r2[sel], c2[sel], k, exact=True)


def test_tile_path_options_agree():
    """K2 variants (queries in TMEM / both operands in shared memory) and the stream plan (overlapped / in line) are
    implementation choices: on exact-arithmetic inputs every combination returns the identical answer."""
    case = make_case(31, 90000, 200, 32, 6, 96, 39, np.uint16, np.uint16, grid=True)
    k = 64
    outs = []
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=32, group=6) as ix:
        configure(ix, 'tile')
        for dv in (1, 0, 2):                                           # 2 = CTA-pair form (tcgen05 cta_group::2, M = 256)
            for ov in (1, 0):
                ix.set_option('dense_variant', dv)
                ix.set_option('overlap', ov)
                ix.set_option('dense_multicast', 2 if ov else 0)       # cluster-of-two multicast also in scratch mode
                ix.set_option('dense_prefetch', ov)                    # TMA L2 prefetch ahead of the demand loads
                ix.set_option('dense_lite', 1 if dv == 1 else 0)       # small-footprint K2 beside K1t (overlapped plan only)
                outs.append(ix.search(case['q_vals'], case['q_idx'], k))
                assert ix.stats()['scan_variant'] == 3, 'tile path not taken'
    sub = dict(case)
    sel = np.r_[0:6, 100:106, 194:200]
    sub['q_vals'], sub['q_idx'] = case['q_vals'][sel], case['q_idx'][sel]
    assert_matches_oracle(sub, outs[0][0][sel], outs[0][1][sel], outs[0][2][sel], k, exact=True)
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])


def test_skewed_index_distribution_zipf():
    """K1t is O(matches); real argmax-over-39 indices are skewed.  Zipf-distributed slice indices
    (big buckets for the hot values, ~3.5x the matches of the uniform recipe) on the tile path and on the row scan."""
    from dhr_b200 import synth
    cv, ci = synth.corpus_numpy('delade_cls_zipf', 0, 40000)
    qv, qi = synth.queries_numpy('delade_cls_zipf', 70)
    case = dict(S=128, G=6, C=768, c_vals=cv, c_idx=ci, q_vals=qv.astype(np.float32), q_idx=qi)
    k = 100
    with GipIndex.from_arrays(cv, ci, n_slices=128, group=6) as ix:
        configure(ix, 'tile')
        s, r, c = ix.search(case['q_vals'], qi, k)
        assert ix.stats()['scan_variant'] == 3
        configure(ix, 'scan1', 8)
        s2, r2, c2 = ix.search(case['q_vals'], qi, k)
    sub = dict(case)
    sel = np.r_[0:5, 60:70]
    sub['q_vals'], sub['q_idx'] = case['q_vals'][sel], qi[sel]
    assert_matches_oracle(sub, s[sel], r[sel], c[sel], k)
    assert_matches_oracle(sub, s2[sel], r2[sel], c2[sel], k)
    assert np.abs(s - s2).max() < 1e-4 and np.mean(r == r2) > 0.99


@pytest.mark.parametrize('shape', [(64, 1, 32), (128, 8, 64), (16, 7, 100), (768, 1, 128), (96, 3, 0), (160, 6, 0)])
@pytest.mark.parametrize('dv', [1, 2])
def test_unmasked_ip_stage_on_tensor_cores(shape, dv):
    """--IP first stage (gip_retrieval.py:139) of a hybrid / lexical index: plain inner product over all columns, run as K2
    column passes (lexical columns from the row-major array, then the dense block; > 768 columns = several passes through
    the scratch).  Grid inputs: bit-exact against the oracle; 200 queries = two query groups (dv 2: CTA-pair form)."""
    S, G, Cd = shape
    case = make_case(300 + S + G, 50000, 200, S, G, Cd, 39, np.uint8, grid=True)
    k = 64
    with GipIndex.from_arrays(case['c_vals'], case['c_idx'], n_slices=S, group=G) as ix:
        ix.set_option('dense_variant', dv)
        s, r, c = ix.search(case['q_vals'], None, k, masked=False)
        assert ix.stats()['scan_variant'] == 4, 'unmasked tile path not taken'
        ix.set_option('tile_mode', 0)
        s1, r1, c1 = ix.search(case['q_vals'][:16], None, k, masked=False)        # row scan K1, same rule
    assert np.array_equal(s[:16], s1) and np.array_equal(r[:16], r1)
    sub = dict(case)
    sel = np.r_[0:5, 125:131, 195:200]
    sub['q_vals'], sub['q_idx'] = case['q_vals'][sel], case['q_idx'][sel]
    assert_matches_oracle(sub, s[sel], r[sel], c[sel], k, masked=False, exact=True)
