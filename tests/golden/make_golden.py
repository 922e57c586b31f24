"""Generate tests/golden/*.npz by running the REAL reference (castorini/dhr @ e236f3d).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture stores the exact inputs handed to the reference
(`retrieval.gip_retrieval.GIP_retrieval` / `IP_retrieval` / `main`) and the
outputs it produced (row lists, float scores, TREC text).  The reference has no
tests or golden vectors of its own, so these fixtures are what
pins the oracle in `oracle/` and, through it, the CUDA path.

Two families of inputs:
  * "grid" cases: every value is a multiple of 1/64 with small magnitude, so all
    products and partial sums are exactly representable in fp32 and the
    reference's scores do not depend on summation order -> score lists can be
    compared bit-exactly; ties are real ties.
  * "gauss" cases: fp16 N(0,1)-like values; compared within the 1e-3 tolerance.
"""
import io
import os
import pickle
import sys
import tempfile
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))
from oracle import refshim  # noqa: E402
from oracle.gip_oracle import make_args  # noqa: E402

ref = refshim.load()


def grid_vals(rng, shape, density, lo=1, hi=128):
    """fp16 values k/64 (k in [lo,hi)), a fraction `density` of entries non-zero."""
    v = rng.integers(lo, hi, size=shape).astype(np.float32) / 64.0
    v *= (rng.random(shape) < density)
    return v.astype(np.float16)


def lexical(rng, n, S, G, R, idx_dtype, density, grid=True):
    """Densified lexical reps: an empty slice has value 0 AND idx 0 (densify_corpus.py:30-45)."""
    if grid:
        vals = grid_vals(rng, (n, S, G), 1.0)
    else:
        vals = np.abs(rng.standard_normal((n, S, G)) * 0.5).astype(np.float16)
    idx = rng.integers(0, R, size=(n, S))
    empty = rng.random((n, S)) >= density
    vals[empty] = 0
    idx[empty] = 0
    return vals.reshape(n, S * G), idx.astype(idx_dtype)


def dense(rng, n, C, grid=True):
    if C == 0:
        return np.zeros((n, 0), np.float16)
    if grid:
        return ((rng.integers(-64, 65, size=(n, C))).astype(np.float32) / 64.0).astype(np.float16)
    return (rng.standard_normal((n, C)) / np.sqrt(C)).astype(np.float16)


def run_gip(case, **argkw):
    """Run ref.GIP_retrieval exactly as main() prepares the tensors on CPU (:274-279,:313-315)."""
    G = case['G']
    qv = torch.from_numpy(case['q_vals'].astype(np.float32))
    cv = torch.from_numpy(case['c_vals'].astype(np.float32))
    qi = torch.from_numpy(np.repeat(case['q_idx'], G, axis=1))     # G>1: one idx per value column
    ci = torch.from_numpy(np.repeat(case['c_idx'], G, axis=1))
    lam = argkw.get('lamda', 1)
    C = case['C']
    if C > 0:
        qv[:, -C:] = lam * qv[:, -C:]                               # :281-283
    args = make_args(emb_dim=case['S'] * G, **argkw)
    qids = list(range(100, 100 + qv.shape[0]))
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        res, sc = ref.GIP_retrieval(qids, qv, qi, cv, ci, args)
    rows = np.array([res[q] for q in qids], dtype=np.int64)
    scores = np.array([sc[q] for q in qids], dtype=np.float64)
    return rows, scores


def run_ip(case, topk):
    qv = torch.from_numpy(case['q_vals'].astype(np.float32))
    cv = torch.from_numpy(case['c_vals'].astype(np.float32))
    qids = list(range(100, 100 + qv.shape[0]))
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        res, sc = ref.IP_retrieval(qids, qv, cv, make_args(topk=topk))
    rows = np.array([res[q] for q in qids], dtype=np.int64)
    scores = np.array([sc[q] for q in qids], dtype=np.float64)
    return rows, scores


def make_case(rng, n, nq, S, G, C, R, c_idx_dtype, q_idx_dtype, c_density, q_density, grid):
    cl, ci = lexical(rng, n, S, G, R, c_idx_dtype, c_density, grid)
    ql, qi = lexical(rng, nq, S, G, R, q_idx_dtype, q_density, grid)
    return dict(S=S, G=G, C=C, c_vals=np.concatenate([cl, dense(rng, n, C, grid)], axis=1),
                c_idx=ci, q_vals=np.concatenate([ql, dense(rng, nq, C, grid)], axis=1), q_idx=qi)


def save(name, case, **outs):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **{k: v for k, v in case.items()}, **outs)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def run_main(tmp, q_triple, c_triple, argv):
    """Run ref.main() on pickle files in `tmp`; returns dict filename -> text of result*.trec."""
    qp, ip = os.path.join(tmp, 'q.pt'), os.path.join(tmp, 'c.index.pt')
    with open(qp, 'wb') as f:
        pickle.dump(q_triple, f, protocol=4)
    with open(ip, 'wb') as f:
        pickle.dump(c_triple, f, protocol=4)
    cwd = os.getcwd()
    os.chdir(tmp)
    old_argv = sys.argv
    try:
        sys.argv = ['gip_retrieval', '--query_emb_path', qp, '--index_path', ip] + argv
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            ref.main()
        out = {}
        for fn in sorted(os.listdir(tmp)):
            if fn.startswith('result') and fn.endswith('.trec'):
                with open(os.path.join(tmp, fn)) as f:
                    out[fn] = f.read()
                os.remove(os.path.join(tmp, fn))
        return out
    finally:
        sys.argv = old_argv
        os.chdir(cwd)


def main():
    torch.set_num_threads(1)                                        # :259, as shipped
    rng = np.random.default_rng(20240917)

    # 1. DeLADE+[CLS] reference-true layout: G=1, uint8 idx in [0,39), dense tail (encode.py:157,166)
    c = make_case(rng, 3000, 8, 64, 1, 32, 39, np.uint8, np.uint8, 0.7, 0.7, grid=True)
    rows, sc = run_gip(c, brute_force=True, topk=100)
    save('delade_g1_u8_grid', c, topk=100, ref_rows=rows, ref_scores=sc)

    # 2. same layout with lamda != 1 (query [CLS] tail scaled in fp32, :281-283)
    c = make_case(rng, 2000, 6, 48, 1, 16, 39, np.uint8, np.uint8, 0.7, 0.7, grid=True)
    rows, sc = run_gip(c, brute_force=True, topk=64, lamda=0.5)
    save('delade_lamda_grid', c, topk=64, lamda=0.5, ref_rows=rows, ref_scores=sc)

    # 3. densified BM25 style: int16 idx both sides, very sparse (range shrunk to 40 so that some rows match at N=4000)
    #    -> many exact ties and zero scores
    c = make_case(rng, 4000, 8, 96, 1, 0, 40, np.int16, np.int16, 0.10, 0.08, grid=True)
    rows, sc = run_gip(c, brute_force=True, topk=80)
    save('bm25_i16_grid', c, topk=80, ref_rows=rows, ref_scores=sc)

    # 4. uniCOIL/SPLADE corpus: int8 idx compared against int16 query idx (densify_corpus.py:33-34, densify_query.py:73)
    c = make_case(rng, 2500, 6, 64, 1, 0, 39, np.int8, np.int16, 0.5, 0.3, grid=True)
    rows, sc = run_gip(c, brute_force=True, topk=50)
    save('unicoil_i8_i16_grid', c, topk=50, ref_rows=rows, ref_scores=sc)

    # 5. grouped generalisation (BASELINE config 2 literal, scaled down): S slices x G=6 values, uint16 idx, dense tail
    c = make_case(rng, 2500, 6, 16, 6, 32, 39, np.uint16, np.uint16, 0.7, 0.7, grid=True)
    rows, sc = run_gip(c, brute_force=True, topk=100)
    save('grouped_g6_u16_grid', c, topk=100, ref_rows=rows, ref_scores=sc)

    # 6. grouped G=3, no dense tail (BASELINE config 3 literal, scaled down)
    c = make_case(rng, 2500, 6, 32, 3, 0, 200, np.uint16, np.uint16, 0.3, 0.25, grid=True)
    rows, sc = run_gip(c, brute_force=True, topk=60)
    save('grouped_g3_u16_grid', c, topk=60, ref_rows=rows, ref_scores=sc)

    # 7. realistic floating-point values (not order-independent): tolerance case
    c = make_case(rng, 3000, 8, 64, 1, 32, 39, np.uint8, np.uint8, 0.7, 0.7, grid=False)
    rows, sc = run_gip(c, brute_force=True, topk=100)
    save('delade_g1_u8_gauss', c, topk=100, ref_rows=rows, ref_scores=sc)

    # 8. dense-only IP_retrieval (Aggretriever, :60-85): full descending argsort, keep k
    c = dict(S=0, G=1, C=64, c_vals=dense(rng, 3000, 64, grid=False), c_idx=np.zeros((3000, 0), np.uint8),
             q_vals=dense(rng, 8, 64, grid=False), q_idx=np.zeros((8, 0), np.uint8))
    rows, sc = run_ip(c, topk=100)
    save('dense_ip_gauss', c, topk=100, ref_rows=rows, ref_scores=sc)
    c = dict(S=0, G=1, C=48, c_vals=dense(rng, 2000, 48, grid=True), c_idx=np.zeros((2000, 0), np.uint8),
             q_vals=dense(rng, 6, 48, grid=True), q_idx=np.zeros((6, 0), np.uint8))
    rows, sc = run_ip(c, topk=2500)                                  # k > N: argsort[:k] just returns N rows
    save('dense_ip_grid_k_gt_n', c, topk=2500, ref_rows=rows, ref_scores=sc)

    # 9. approximate modes (:128-156): theta pruning, --IP first stage, exact rerank of agip_topk
    c = make_case(rng, 3000, 6, 64, 1, 32, 39, np.uint8, np.uint8, 0.7, 0.7, grid=True)
    outs = {}
    for tag, kw in [('theta', dict(theta=0.8)), ('theta_rerank', dict(theta=0.8, rerank=True, agip_topk=400)),
                    ('ip', dict(theta=0.8, IP=True)), ('ip_rerank', dict(theta=0.8, IP=True, rerank=True, agip_topk=400))]:
        rows, sc = run_gip(c, topk=50, **kw)
        outs['ref_rows_' + tag], outs['ref_scores_' + tag] = rows, sc
    save('delade_approx_grid', c, topk=50, theta=0.8, agip_topk=400, **outs)

    # 10. main(): pickle files in, TREC text out; single shard and --total_shrad 2
    with tempfile.TemporaryDirectory() as tmp:
        c = make_case(rng, 400, 5, 32, 1, 16, 39, np.uint8, np.uint8, 0.7, 0.7, grid=True)
        docids = [str(7000 + i) for i in range(400)]
        qids = ['7003', 'q1', 'q2', '7150', 'q4']                   # two qids collide with docids (:340 skip rule)
        q_triple = [c['q_vals'], c['q_idx'], qids]
        c_triple = [c['c_vals'], c['c_idx'], docids]
        outs = {}
        t = run_main(tmp, q_triple, c_triple, ['--emb_dim', '32', '--brute_force', '--topk', '20', '--lamda', '0.5',
                                                 '--run_name', 'golden'])
        outs['trec_single'] = np.array(t['result.trec'])
        for sh in (0, 1, 2):
            t = run_main(tmp, q_triple, c_triple, ['--emb_dim', '32', '--brute_force', '--topk', '20', '--lamda', '0.5',
                                                     '--total_shrad', '3', '--shrad', str(sh), '--run_name', 'golden'])
            outs['trec_shard%d' % sh] = np.array(t['result%d.trec' % sh])
        save('main_trec_grid', c, topk=20, lamda=0.5, docids=np.array(docids), qids=np.array(qids), **outs)

        # dense-only through main(): idx entries are None / 0 (encode.py:149-153, index.py:40-43)
        cd = dict(S=0, G=1, C=32, c_vals=dense(rng, 300, 32, grid=True), c_idx=np.zeros((300, 0), np.uint8),
                  q_vals=dense(rng, 4, 32, grid=True), q_idx=np.zeros((4, 0), np.uint8))
        docids = [str(i) for i in range(300)]
        qids = ['a', 'b', 'c', 'd']
        t = run_main(tmp, [cd['q_vals'], None, qids], [cd['c_vals'], 0, docids], ['--topk', '15', '--run_name', 'golden'])
        save('main_trec_dense_grid', cd, topk=15, docids=np.array(docids), qids=np.array(qids),
             trec_single=np.array(t['result.trec']))


def densify_golden():
    """tevatron/DHR/utils.py:densify on a small vocabulary (same structure as 30522 = 570 + 39 * 768)."""
    sys.path.insert(0, refshim.REFERENCE_ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_dhr_utils', os.path.join(refshim.REFERENCE_ROOT, 'tevatron', 'DHR', 'utils.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(7)
    x = np.log1p(np.maximum(rng.standard_normal((6, 30 + 39 * 64)).astype(np.float32), 0))      # relu-log style reps, many exact zeros/ties
    vals, idx = mod.densify(torch.from_numpy(x), dims=64, remove_dims=30)
    np.savez_compressed(os.path.join(HERE, 'densify_op.npz'), x=x, dims=64, remove_dims=30,
                        ref_vals=vals.numpy().astype(np.float16), ref_idx=idx.numpy().astype(np.uint8))
    print('wrote densify_op.npz')


def wide_golden():
    """Index ranges beyond 8 bits (densified BM25 / DeepImpact: idx < 3466, densify_corpus.py:31-32): the CUDA tile path
    switches to 16-bit codes there.  Indices are drawn from a pool of 60 values spread over [0, 3466) so that queries and
    passages do collide at this small N.  Own rng: the fixtures written by main() stay byte-identical."""
    torch.set_num_threads(1)
    rng = np.random.default_rng(20261017)
    pool = np.sort(rng.choice(3466, size=60, replace=False))

    def remap(case):
        for key in ('c_idx', 'q_idx'):
            idx = case[key]
            nz = np.any(case[key[0] + '_vals'][:, :case['S'] * case['G']].reshape(idx.shape[0], case['S'], case['G']) != 0, axis=2)
            new = pool[idx.astype(np.int64) % 60]
            new[~nz] = 0                                               # empty slice: value 0 and idx 0
            case[key] = new.astype(idx.dtype)
        return case

    # reference-true densified BM25: G = 1, int16 idx on both sides
    c = remap(make_case(rng, 4000, 8, 64, 1, 0, 60, np.int16, np.int16, 0.30, 0.25, grid=True))
    rows, sc = run_gip(c, brute_force=True, topk=80)
    save('bm25_wide_i16_grid', c, topk=80, ref_rows=rows, ref_scores=sc)
    # BASELINE config 3 literal, scaled down: 3 values per slice, uint16 idx
    c = remap(make_case(rng, 3000, 6, 32, 3, 0, 60, np.uint16, np.uint16, 0.30, 0.30, grid=True))
    rows, sc = run_gip(c, brute_force=True, topk=60)
    save('grouped_g3_wide_u16_grid', c, topk=60, ref_rows=rows, ref_scores=sc)


def pq_golden():
    """PQ_IP_retrieval (:167-231): the faiss IndexPQ first stage is third-party and absent here, so the reference's own
    rerank half (:205-215) and its no-rerank branch (:218-222) are executed on candidate lists served by a stub index
    (`faiss.read_index` -> object with .search()).  Candidates: the exact-IP top 200 of each query plus 200 random other
    rows, shuffled -- unique per query, like an IndexPQ result.  Own rng: other fixtures stay byte-identical."""
    torch.set_num_threads(1)
    rng = np.random.default_rng(20261018)
    c = make_case(rng, 3000, 6, 64, 1, 32, 39, np.uint8, np.uint8, 0.7, 0.7, grid=True)
    n, nq, M, k = 3000, 6, 400, 50
    ip = c['q_vals'].astype(np.float32) @ c['c_vals'].astype(np.float32).T
    cands = np.zeros((nq, M), np.int64)
    for i in range(nq):
        top = np.argsort(-ip[i], kind='stable')[:200]
        rest = np.setdiff1d(np.arange(n), top)
        cands[i] = rng.permutation(np.concatenate([top, rng.choice(rest, size=M - 200, replace=False)]))
    cscores = np.take_along_axis(ip, cands, axis=1).astype(np.float32)

    class StubIndex:                       # what faiss.read_index returns: .search(x, k) -> (D, I), batches served in order
        def __init__(self):
            self.cursor = 0

        def search(self, x, kk):
            b = x.shape[0]
            out = cscores[self.cursor:self.cursor + b, :kk], cands[self.cursor:self.cursor + b, :kk]
            self.cursor += b
            return out

    G = c['G']
    qv = torch.from_numpy(c['q_vals'].astype(np.float32))
    cv = torch.from_numpy(c['c_vals'].astype(np.float32))
    qi = torch.from_numpy(c['q_idx'])
    ci = torch.from_numpy(c['c_idx'])
    qids = list(range(100, 100 + nq))
    outs = {}
    for tag, rerank in (('rerank', True), ('norerank', False)):
        ref.faiss.read_index = lambda path: StubIndex()
        args = make_args(emb_dim=c['S'] * G, topk=k, agip_topk=M, rerank=rerank, batch=4, faiss_pq_index_path='stub')
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            res, sc = ref.PQ_IP_retrieval(qids, qv, qi, cv, ci, args)
        outs['ref_rows_' + tag] = np.array([[int(x) for x in res[q]] for q in qids], dtype=np.int64)
        outs['ref_scores_' + tag] = np.array([[float(x) for x in sc[q]] for q in qids], dtype=np.float64)
    save('pq_rerank_grid', c, topk=k, agip_topk=M, candidates=cands, candidate_scores=cscores, **outs)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'pq':
        pq_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == 'densify':
        densify_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == 'wide':
        wide_golden()
    else:
        main()
        densify_golden()
        wide_golden()
        pq_golden()
