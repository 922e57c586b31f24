"""Shared helpers for the parity tests (the oracle is the checker, never the thing under test)."""
import numpy as np

from oracle import gip_oracle as go
from oracle import c_oracle as co


def make_case(seed, n, nq, S, G, C, R, c_idx_dtype=np.uint8, q_idx_dtype=None, c_density=0.7, q_density=0.7,
              grid=False, q_fp32_noise=False):
    """Random encoded corpus/queries in the reference's layout (empty slice = value 0 and idx 0)."""
    rng = np.random.default_rng(seed)
    q_idx_dtype = q_idx_dtype or c_idx_dtype

    def lex(m, density, dtype):
        if S == 0:
            return np.zeros((m, 0), np.float16), np.zeros((m, 0), dtype)
        if grid:
            v = rng.integers(1, 128, size=(m, S, G)).astype(np.float32) / 64.0
        else:
            v = np.abs(rng.standard_normal((m, S, G)) * 0.5).astype(np.float32)
        ii = rng.integers(0, R, size=(m, S))
        empty = rng.random((m, S)) >= density
        v[empty] = 0
        ii[empty] = 0
        return v.reshape(m, S * G).astype(np.float16), ii.astype(dtype)

    def dense(m):
        if C == 0:
            return np.zeros((m, 0), np.float16)
        if grid:
            return (rng.integers(-64, 65, size=(m, C)).astype(np.float32) / 64.0).astype(np.float16)
        return (rng.standard_normal((m, C)) / np.sqrt(max(C, 1))).astype(np.float16)

    cl, ci = lex(n, c_density, c_idx_dtype)
    ql, qi = lex(nq, q_density, q_idx_dtype)
    c_vals = np.concatenate([cl, dense(n)], axis=1)
    q_vals = np.concatenate([ql, dense(nq)], axis=1).astype(np.float32)
    if q_fp32_noise:   # make queries NOT representable in fp16 (exercises the fp32-query kernels)
        q_vals = (q_vals * np.float32(1.0009765625 + 1e-4)).astype(np.float32)
    return dict(S=S, G=G, C=C, c_vals=c_vals, c_idx=ci, q_vals=q_vals, q_idx=qi)


def assert_matches_oracle(case, scores, rows, counts, k, masked=True, lamda=1.0, atol=1e-3, row_offset=0, exact=False):
    """Compare a [Q,k] result with the exact fp64 oracle (C implementation)."""
    q = case['q_vals'].astype(np.float32).copy()
    C = case['C']
    if C > 0 and lamda != 1.0:
        q[:, -C:] = np.float32(lamda) * q[:, -C:]
    n = case['c_vals'].shape[0]
    kk = min(k, n)
    assert scores.shape == (q.shape[0], k) and rows.shape == (q.shape[0], k)
    assert np.all(counts == kk)
    if kk < k:
        assert np.all(rows[:, kk:] == -1) and np.all(np.isneginf(scores[:, kk:]))
    for i in range(q.shape[0]):
        ex = co.scores(case['c_vals'], case['c_idx'], q[i], case['q_idx'][i] if case['S'] > 0 else None,
                       case['S'], case['G'], masked=masked)
        r = rows[i, :kk] - row_offset
        if exact:   # grid inputs: fp32 arithmetic is exact -> bit-identical scores and deterministic order
            er, es = go.topk_desc(ex, kk)
            assert np.array_equal(r, er), 'query %d: rows differ from (score desc, row asc) oracle order' % i
            assert np.array_equal(scores[i, :kk].astype(np.float64), es), 'query %d: scores not bit-exact' % i
        else:
            msg = go.check_topk_against_exact(r, scores[i, :kk], ex, k, atol=atol)
            assert msg is None, 'query %d: %s' % (i, msg)


def tie_groups_equal(rows_a, rows_b, scores):
    n = len(scores)
    i = 0
    while i < n:
        j = i
        while j + 1 < n and scores[j + 1] == scores[i]:
            j += 1
        if j < n - 1 and set(rows_a[i:j + 1].tolist()) != set(rows_b[i:j + 1].tolist()):
            return False
        i = j + 1
    return True
