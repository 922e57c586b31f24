"""CPU oracle for the GIP retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy / torch, the algorithm of castorini/dhr
``retrieval/gip_retrieval.py`` (reference @ e236f3d).  Nothing under
``dhr_b200/`` imports this module; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` arm do.

Parity: pinned by execution against the real reference (see
``tests/golden/make_golden.py`` and ``tests/test_oracle_golden.py``).

Grouped inner product, as the reference computes it (gip_retrieval.py:110-125)::

    score[p] = sum_{s<S} [q_idx[s] == p_idx[s]] * sum_{g<G} q_val[s*G+g] * p_val[s*G+g]
             + sum_{c<C} q_val[D+c] * p_val[D+c]            D = S*G, W = D + C

The reference is the G == 1 case; for G > 1 feed it ``np.repeat(idx, G, axis=1)``.
"""
from __future__ import annotations

import time
from types import SimpleNamespace

import numpy as np

try:  # torch is only needed by the timed port
    import torch
except Exception:  # pragma: no cover
    torch = None


# --------------------------------------------------------------------------
# exact scorer (fp64) and deterministic top-k
# --------------------------------------------------------------------------
def gip_scores_f64(q_vals, q_idx, c_vals, c_idx, n_slices, group=1, masked=True):
    """Exact fp64 GIP score matrix [Q, N] for small inputs.

    Follows gip_retrieval.py:110-120: the equality mask is applied to the first
    ``n_slices*group`` columns (one index per slice, shared by ``group``
    values), the remaining columns are the always-matching dense tail (the
    reference pads both index arrays with the constant 1, :110-113).
    ``masked=False`` is the ``--IP`` branch (:139): plain inner product.
    Integer equality is on the integer value after promotion (torch promotes
    int8 corpus vs int16 query, :119).
    """
    q = np.asarray(q_vals, dtype=np.float64)
    c = np.asarray(c_vals, dtype=np.float64)
    Q, W = q.shape
    N = c.shape[0]
    D = n_slices * group
    out = np.zeros((Q, N), dtype=np.float64)
    if D > 0:
        if masked and q_idx is not None:
            qi = np.asarray(q_idx).astype(np.int64)
            ci = np.asarray(c_idx).astype(np.int64)
            for i in range(Q):
                m = (ci == qi[i][None, :])                       # [N, S]
                if group > 1:
                    m = np.repeat(m, group, axis=1)              # [N, D]
                out[i] = (np.where(m, c[:, :D], 0.0) * q[i, :D][None, :]).sum(axis=1)
        else:
            out += q[:, :D] @ c[:, :D].T
    if W > D:
        out += q[:, D:] @ c[:, D:].T
    return out


def topk_desc(scores, k):
    """Deterministic top-k: (score desc, row asc).  Returns (rows[k], scores[k]).

    The reference uses torch.topk (:123) / argsort(descending) (:75) whose tie
    order is unspecified; this is the tie rule the CUDA path implements.
    """
    s = np.asarray(scores)
    k = min(k, s.shape[0])
    order = np.lexsort((np.arange(s.shape[0]), -s))[:k]
    return order.astype(np.int64), s[order]


def search_f64(q_vals, q_idx, c_vals, c_idx, n_slices, group, k, masked=True):
    """[Q,k] rows (int64) and exact scores (fp64) with the deterministic tie rule."""
    sc = gip_scores_f64(q_vals, q_idx, c_vals, c_idx, n_slices, group, masked)
    rows = np.empty((sc.shape[0], min(k, sc.shape[1])), dtype=np.int64)
    vals = np.empty(rows.shape, dtype=np.float64)
    for i in range(sc.shape[0]):
        rows[i], vals[i] = topk_desc(sc[i], k)
    return rows, vals


# --------------------------------------------------------------------------
# range sharding (gip_retrieval.py:292-306) and shard merge (merge.result.py:20-43)
# --------------------------------------------------------------------------
def shard_bounds(n_docs, total_shards, shard):
    """Rows [lo, hi) of shard ``shard``: floor(N/T) rows each, remainder to the last."""
    per = n_docs // total_shards
    lo = per * shard
    hi = n_docs if shard == total_shards - 1 else per * (shard + 1)
    return lo, hi


def merge_topk(shard_scores, shard_rows, k):
    """Merge per-shard top-k lists ([P,Q,k] each, global row ids) into [Q,k].

    merge.result.py:39 concatenates all shards' (docid, score) per query and
    keeps the k best by score; ties there follow ``argsort()[::-1]`` (arbitrary
    w.r.t. the single-shard run), here they follow (score desc, row asc) so an
    8-shard run equals the 1-shard run exactly.
    """
    s = np.asarray(shard_scores)
    r = np.asarray(shard_rows)
    P, Q, kk = s.shape
    out_s = np.empty((Q, min(k, P * kk)), dtype=s.dtype)
    out_r = np.empty(out_s.shape, dtype=np.int64)
    for q in range(Q):
        cs = s[:, q, :].reshape(-1)
        cr = r[:, q, :].reshape(-1)
        keep = cr >= 0
        cs, cr = cs[keep], cr[keep]
        order = np.lexsort((cr, -cs))[:k]
        n = order.shape[0]
        out_s[q, :n], out_r[q, :n] = cs[order], cr[order]
        out_s[q, n:], out_r[q, n:] = -np.inf, -1
    return out_s, out_r


# --------------------------------------------------------------------------
# torch-op port of the reference loops (the timed CPU baseline, kind="port")
# --------------------------------------------------------------------------
def GIP_retrieval_port(qids, query_embs, query_arg_idxs, corpus_embs, corpus_arg_idxs, args,
                       verbose=False):
    """Same operator sequence as gip_retrieval.py:88-165 on torch tensors.

    Exact branch (:117-126): eq-mask -> multiply -> row-dot -> topk.
    Approximate branch (:128-156): theta pruning of query dims, partial GIP over
    the kept columns or unmasked IP (--IP), optional exact rerank of agip_topk.
    Returns (results, scores) dicts exactly like the reference.
    """
    theta = 0 if args.brute_force else args.theta
    tail = query_embs.shape[1] - args.emb_dim
    q_idx, c_idx = query_arg_idxs, corpus_arg_idxs
    if tail > 0:                                              # :110-113
        q_idx = torch.nn.functional.pad(q_idx, (0, tail), value=1)
        c_idx = torch.nn.functional.pad(c_idx, (0, tail), value=1)
    results, scores_out = {}, {}
    t0 = time.time()
    for n, (qv, qi) in enumerate(zip(query_embs, q_idx)):
        if theta == 0:
            gated = (c_idx == qi) * corpus_embs               # :119
            sc = torch.mv(gated, qv)                          # :120 einsum('ij,j->i')
            del gated
            top = torch.topk(sc, args.topk, dim=0).indices    # :123
            rows, vals = top, sc[top]
        else:
            keep_n = int((qv > theta).sum())                  # :130
            keep = torch.topk(qv, keep_n, dim=0).indices.tolist()   # :131
            if not args.IP:                                   # :135-136
                gated = (c_idx[:, keep] == qi[keep]) * corpus_embs[:, keep]
                part = torch.mv(gated, qv[keep])
            else:                                             # :139
                part = torch.mv(corpus_embs, qv)
            if args.rerank:                                   # :142-150
                cand = torch.topk(part, args.agip_topk, dim=0).indices
                gated = (c_idx[cand, :] == qi) * corpus_embs[cand]
                sc = torch.mv(gated, qv)
                top = torch.topk(sc, args.topk, dim=0).indices
                rows, vals = cand[top], sc[top]
            else:                                             # :155-156
                rows = torch.topk(part, args.topk, dim=0).indices
                vals = part[rows]
        scores_out[qids[n]] = vals.tolist()
        results[qids[n]] = rows.tolist()
    if verbose:
        print('Retrieving {} queries ({:0.3f} s/query)'.format(
            len(query_embs), (time.time() - t0) / max(1, len(query_embs))))
    return results, scores_out


def PQ_rerank_port(qids, query_embs, query_arg_idxs, corpus_embs, corpus_arg_idxs, candidates, candidate_scores, args):
    """gip_retrieval.py:167-231 given the first-stage lists (faiss IndexPQ.search is third-party, not restated):
    rerank half (:205-215) = exact GIP over the candidates, top args.topk; otherwise (:218-222) the lists as they are."""
    tail = query_embs.shape[1] - query_arg_idxs.shape[1]
    q_idx, c_idx = query_arg_idxs, corpus_arg_idxs
    if tail > 0:                                              # :179-182
        q_idx = torch.nn.functional.pad(q_idx, (0, tail), value=1)
        c_idx = torch.nn.functional.pad(c_idx, (0, tail), value=1)
    results, scores_out = {}, {}
    for n, (qv, qi) in enumerate(zip(query_embs, q_idx)):
        cand = torch.as_tensor(np.asarray(candidates[n]), dtype=torch.long)
        if args.rerank:
            gated = (c_idx[cand, :] == qi) * corpus_embs[cand]      # :206
            sc = torch.mv(gated, qv)                                # :207
            top = torch.topk(sc, args.topk, dim=0).indices          # :209
            results[qids[n]] = cand[top].tolist()
            scores_out[qids[n]] = sc[top].tolist()
        else:
            results[qids[n]] = cand[:args.topk].tolist()
            scores_out[qids[n]] = np.asarray(candidate_scores[n])[:args.topk].tolist()
    return results, scores_out


def IP_retrieval_port(qids, query_embs, corpus_embs, args):
    """gip_retrieval.py:60-85: row-dot then full descending argsort, keep topk."""
    results, scores_out = {}, {}
    for n, qv in enumerate(query_embs):
        sc = torch.mv(corpus_embs, qv)                        # :74
        rows = torch.argsort(sc, descending=True)[:args.topk]  # :75
        scores_out[qids[n]] = sc[rows].tolist()
        results[qids[n]] = rows.tolist()
    return results, scores_out


def densify_port(lexical_reps, dims=768, remove_dims=570):
    """numpy restatement of tevatron/DHR/utils.py:5-22 + the storage casts of tevatron/driver/encode.py:156-157:
    drop the first remove_dims ids, view (B, R, dims), max over R; fp16 values, uint8 argmax (first maximum)."""
    x = np.asarray(lexical_reps, dtype=np.float32)
    b = x.shape[0]
    if (x.shape[1] - remove_dims) % dims != 0:
        raise ValueError('Input lexical representation cannot be densified, please fix dims or remove_dims')
    v = x[:, remove_dims:].reshape(b, -1, dims)
    return v.max(axis=1).astype(np.float16), v.argmax(axis=1).astype(np.uint8)


def make_args(**kw):
    """Namespace with the defaults of gip_retrieval.py:234-253."""
    d = dict(emb_dim=768, theta=0.1, topk=1000, agip_topk=10000, combine_cls=False, IP=False,
             PQIP=False, batch=1, brute_force=False, use_gpu=False, rerank=False, lamda=1,
             total_shrad=1, shrad=0, run_name='h2oloo', faiss_pq_index_path=None)
    d.update(kw)
    return SimpleNamespace(**d)


def trec_lines(results, scores, docids, run_name):
    """gip_retrieval.py:333-341: skip docid == qid, rank = position + 1 (not renumbered)."""
    out = []
    for qid in results:
        for rank, row in enumerate(results[qid]):
            docid = docids[row]
            if docid != qid:
                out.append('{} Q0 {} {} {} {}\n'.format(qid, docid, rank + 1, scores[qid][rank], run_name))
    return out


# --------------------------------------------------------------------------
# comparison helpers shared by the parity tests
# --------------------------------------------------------------------------
def check_topk_against_exact(rows, vals, exact_scores, k, atol=1e-3, tie_eps=None):
    """Check one query's result list against the exact fp64 score vector.

    * every returned score is within ``atol`` of the exact score of that row
      (north_star: scores within 1e-3),
    * the list is non-increasing and has no duplicate rows,
    * ranks are exact outside tie groups: any position where the returned row
      differs from the deterministic exact ranking must sit in a group of rows
      whose exact scores differ by <= tie_eps (fp32 reorder noise / true ties),
    * nothing better than the k-th exact score (beyond tie_eps) is missing.
    Returns None if fine, else a string describing the first violation.
    """
    rows = np.asarray(rows, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    ex = np.asarray(exact_scores, dtype=np.float64)
    n = ex.shape[0]
    kk = min(k, n)
    if rows.shape[0] != kk:
        return f'length {rows.shape[0]} != {kk}'
    if tie_eps is None:
        tie_eps = 4e-6 * max(1.0, float(np.abs(ex).max()))
    if len(set(rows.tolist())) != kk:
        return 'duplicate rows'
    if rows.min() < 0 or rows.max() >= n:
        return 'row out of range'
    if np.any(np.diff(vals) > 0):
        return 'scores not non-increasing'
    err = np.abs(vals - ex[rows])
    if err.max() > atol:
        return f'score error {err.max():.3e} > {atol}'
    ref_rows, ref_vals = topk_desc(ex, kk)
    for pos in np.nonzero(rows != ref_rows)[0]:
        if abs(ex[rows[pos]] - ref_vals[pos]) > tie_eps:
            return (f'rank {pos}: row {rows[pos]} (exact {ex[rows[pos]]:.9g}) vs oracle row '
                    f'{ref_rows[pos]} (exact {ref_vals[pos]:.9g}) differ beyond tie_eps={tie_eps:.2e}')
    return None
