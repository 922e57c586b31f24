"""Drop-in mirror of castorini/dhr ``retrieval/gip_retrieval.py`` on the B200 CUDA path.

Same function names, argument meaning, return types and command line as the reference
(``GIP_retrieval`` :88-165, ``IP_retrieval`` :60-85, ``main`` :233-344); the per-query
torch loop is replaced by the HBM-resident index and fused scan/top-k kernels behind the C ABI
(include/dhr_b200.h).  ``python -m dhr_b200.gip_retrieval`` accepts the reference's flags verbatim
(misspellings included: --total_shrad, --shrad, --lamda) and reads / writes the same files.

Differences, all deliberate:
  * ties are ordered (score desc, row asc); torch.topk's tie order is unspecified,
  * ``PQ_IP_retrieval`` (:167-231): the faiss IndexPQ first stage is third-party; its candidate lists are taken from
    ``candidates=`` / ``first_stage=`` / a candidate file (or faiss itself when installed), the exact GIP rerank half
    (:205-215) runs on the device,
  * ``--use_gpu`` is accepted and ignored (this implementation always runs on the GPU); the
    extra flag ``--device`` selects the CUDA device (the reference hard-wires 0, :261).
"""
from __future__ import annotations

import argparse
import ctypes
import os
import pickle
import time

import numpy as np

from . import _cabi
from .index import GipIndex

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _as_lists(scores, rows, counts, qids, local_offset=0):
    """[Q,k] arrays -> the reference's two dicts (qid -> list[int], qid -> list[float])."""
    if torch is not None and isinstance(scores, torch.Tensor):
        scores, rows, counts = scores.cpu().numpy(), rows.cpu().numpy(), counts.cpu().numpy()
    all_results, all_scores = {}, {}
    for i, qid in enumerate(qids):
        n = int(counts[i])
        all_scores[qid] = scores[i, :n].tolist()
        all_results[qid] = (rows[i, :n] - local_offset).tolist()
    return all_results, all_scores


def _open_index(corpus_embs, corpus_arg_idxs, emb_dim, device=0):
    """corpus_embs may already be a GipIndex (extension: lets callers keep the corpus resident)."""
    if isinstance(corpus_embs, GipIndex):
        return corpus_embs, False
    if corpus_arg_idxs is None:
        return GipIndex.from_arrays(corpus_embs, None, device=device), True
    return GipIndex.from_arrays(corpus_embs, corpus_arg_idxs, n_slices=emb_dim, group=1, device=device), True


def IP_retrieval(qids, query_embs, corpus_embs, args):
    """Brute-force inner-product search (gip_retrieval.py:60-85): top ``args.topk`` rows per query.

    Like the reference's ``argsort(...)[:topk]`` this never fails for topk > N; it returns N rows."""
    description = 'Brute force IP search'
    index, owned = _open_index(corpus_embs, None, 0, getattr(args, 'device', 0))
    try:
        start_time = time.time()
        k = max(1, min(int(args.topk), max(1, len(index))))
        scores, rows, counts = index.search(query_embs, None, k)
        all_results, all_scores = _as_lists(scores, rows, counts, qids, index.row_offset)
        time_per_query = (time.time() - start_time) / max(1, len(qids))
        print('Retrieving {} queries ({:0.3f} s/query), average number of index use {}'.format(
            len(qids), time_per_query, 0.0))
    finally:
        if owned:
            index.close()
    del description
    return all_results, all_scores


def _prune_query(query_embs, theta):
    """theta pruning of :130-131: keep the dimensions whose query value exceeds theta (all columns,
    [CLS] tail included), zero the rest -- identical to restricting the sum to the kept columns."""
    if torch is not None and isinstance(query_embs, torch.Tensor):
        return torch.where(query_embs > theta, query_embs, torch.zeros_like(query_embs))
    q = np.asarray(query_embs)
    return np.where(q > theta, q, np.zeros_like(q))


def GIP_retrieval(qids, query_embs, query_arg_idxs, corpus_embs, corpus_arg_idxs, args):
    """Grouped-inner-product search (gip_retrieval.py:88-165).

    args fields read: brute_force, theta, IP, rerank, emb_dim, topk, agip_topk (args.theta is set to 0
    under --brute_force exactly like :90).  Returns (all_results, all_scores): qid -> list of row
    indices within the given corpus, qid -> list of fp32 scores, both descending by score."""
    if args.brute_force:
        args.theta = 0
    index, owned = _open_index(corpus_embs, corpus_arg_idxs, args.emb_dim, getattr(args, 'device', 0))
    try:
        n = len(index)
        start_time = time.time()
        total_num_idx = 0
        if args.theta == 0:
            if args.topk > n:      # torch.topk raises here (:123)
                raise RuntimeError('selected index k out of range')
            total_num_idx = args.emb_dim * len(qids)
            scores, rows, counts = index.search(query_embs, query_arg_idxs, args.topk)
        else:
            if not args.IP:
                first = index.search(_prune_query(query_embs, args.theta), query_arg_idxs,
                                     args.agip_topk if args.rerank else args.topk)      # :135-136
            else:
                first = index.search(query_embs, None, args.agip_topk if args.rerank else args.topk,
                                     masked=False)                                       # :139
            need = args.agip_topk if args.rerank else args.topk
            if need > n:
                raise RuntimeError('selected index k out of range')
            if args.rerank:                                                              # :142-150
                if args.topk > args.agip_topk:
                    raise RuntimeError('selected index k out of range')
                cand = first[1] - index.row_offset
                scores, rows, counts = index.rerank(query_embs, query_arg_idxs, cand, args.topk)
            else:                                                                        # :155-156
                scores, rows, counts = first
        all_results, all_scores = _as_lists(scores, rows, counts, qids, index.row_offset)
        time_per_query = (time.time() - start_time) / max(1, len(qids))
        print('Retrieving {} queries ({:0.3f} s/query), average number of index use {}'.format(
            len(qids), time_per_query, total_num_idx / max(1, len(qids))))
    finally:
        if owned:
            index.close()
    return all_results, all_scores


def _first_stage_candidates(query_embs, args, candidates, candidate_scores, first_stage):
    """Candidate lists [Q, agip_topk] of the PQ first stage (gip_retrieval.py:202).  faiss is third-party and not part of
    this path, so the lists come from (in this order) `candidates` (+ `candidate_scores`), a `first_stage(batch) -> (D, I)`
    callable with faiss' IndexPQ.search signature, a candidate file given as --faiss_pq_index_path (.npy [Q, M] or .npz
    with 'candidates' and optionally 'scores'), or a faiss index file when faiss is importable."""
    if candidates is not None:
        return np.asarray(candidates), None if candidate_scores is None else np.asarray(candidate_scores)
    path = getattr(args, 'faiss_pq_index_path', None)
    if first_stage is None:
        assert path is not None, 'you do not spesify your PQ index through --faiss_pq_index_path'
        if str(path).endswith('.npy'):
            return np.load(path), None
        if str(path).endswith('.npz'):
            with np.load(path) as z:
                return z['candidates'], (z['scores'] if 'scores' in z.files else None)
        try:
            import faiss
        except Exception as e:
            raise RuntimeError('PQ_IP_retrieval: faiss is not installed; pass candidates=/first_stage= or point '
                               '--faiss_pq_index_path at a .npy/.npz candidate file') from e
        print('Load PQ index ...')
        first_stage = faiss.read_index(path).search
    q = query_embs.cpu().numpy() if torch is not None and isinstance(query_embs, torch.Tensor) else np.asarray(query_embs)
    q = np.ascontiguousarray(q, dtype=np.float32)
    batch = max(1, int(getattr(args, 'batch', 1)))
    D, I = [], []
    for b in range(0, q.shape[0], batch):                            # :186-202 batching
        d, i = first_stage(q[b:b + batch], args.agip_topk)
        D.append(np.asarray(d)); I.append(np.asarray(i))
    return np.concatenate(I), np.concatenate(D)


def PQ_IP_retrieval(qids, query_embs, query_arg_idxs, corpus_embs, corpus_arg_idxs, args, candidates=None,
                    candidate_scores=None, first_stage=None):
    """Product-quantised first stage + exact GIP rerank (gip_retrieval.py:167-231).

    The rerank half (:205-215) -- exact GIP of every query over its agip_topk candidates, top args.topk -- runs on the
    device (dhr_rerank).  Without --rerank the first-stage lists are returned as they are (:218-222), which needs
    candidate scores.  Candidate ids are rows of the given corpus, unique per query; ids < 0 (faiss' "no result") are
    skipped."""
    cand, cand_scores = _first_stage_candidates(query_embs, args, candidates, candidate_scores, first_stage)
    if cand.ndim != 2 or cand.shape[0] != len(qids):
        raise ValueError('candidates must be [n_queries, agip_topk], got %s' % (cand.shape,))
    start_time = time.time()
    if not args.rerank:
        if cand_scores is None:
            raise ValueError('PQ_IP_retrieval without --rerank returns the first-stage scores: candidate scores are required')
        all_results = {qid: cand[i, :args.topk].tolist() for i, qid in enumerate(qids)}
        all_scores = {qid: np.asarray(cand_scores)[i, :args.topk].tolist() for i, qid in enumerate(qids)}
    else:
        index, owned = _open_index(corpus_embs, corpus_arg_idxs, args.emb_dim, getattr(args, 'device', 0))
        try:
            if args.topk > cand.shape[1]:                             # torch.topk raises here (:209)
                raise RuntimeError('selected index k out of range')
            scores, rows, counts = index.rerank(query_embs, query_arg_idxs, cand.astype(np.int64), args.topk)
            all_results, all_scores = _as_lists(scores, rows, counts, qids, index.row_offset)
        finally:
            if owned:
                index.close()
    time_per_query = (time.time() - start_time) / max(1, len(qids))
    print('Retrieving {} queries ({:0.3f} s/query)'.format(len(qids), time_per_query))
    return all_results, all_scores


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--query_emb_path", type=str, required=True)
    parser.add_argument("--index_path", type=str, required=True)
    parser.add_argument("--faiss_pq_index_path", type=str, default=None)
    parser.add_argument("--emb_dim", type=int, default=768, help='DLR dimension')
    parser.add_argument("--theta", type=float, default=0.1)
    parser.add_argument("--topk", type=int, default=1000)
    parser.add_argument("--agip_topk", type=int, default=10000)
    parser.add_argument("--combine_cls", action='store_true')
    parser.add_argument("--IP", action='store_true')
    parser.add_argument("--PQIP", action='store_true')
    parser.add_argument("--batch", type=int, default=1)
    parser.add_argument("--brute_force", action='store_true')
    parser.add_argument("--use_gpu", action='store_true')
    parser.add_argument("--rerank", action='store_true')
    parser.add_argument("--lamda", type=float, default=1, help='weight for [CSL] for concatenation')
    parser.add_argument("--total_shrad", type=int, default=1)
    parser.add_argument("--shrad", type=int, default=0)
    parser.add_argument("--run_name", type=str, default='h2oloo')
    parser.add_argument("--device", type=int, default=0, help='CUDA device (extension; the reference hard-wires 0)')
    return parser


def shard_bounds(n_docs, total_shrad, shrad):
    """gip_retrieval.py:292-306: floor(N/T) rows per shard, the last shard takes the remainder."""
    per = n_docs // total_shrad
    lo = per * shrad
    hi = n_docs if shrad == total_shrad - 1 else per * (shrad + 1)
    return lo, hi


def _id_table(ids):
    """ids (list / array of ints or of strs) -> (kind, int64 array | None, packed utf-8 bytes | None, offsets | None)."""
    if isinstance(ids, np.ndarray) and ids.dtype.kind in 'iu':
        return 'int', np.ascontiguousarray(ids, dtype=np.int64), None, None
    ids = list(ids)
    if all(isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_)) for x in ids):
        try:
            return 'int', np.asarray(ids, dtype=np.int64), None, None
        except OverflowError:
            return None, None, None, None
    if all(isinstance(x, str) for x in ids):
        enc = [x.encode('utf-8') for x in ids]
        off = np.zeros(len(enc) + 1, np.int64)
        np.cumsum([len(b) for b in enc], out=off[1:])
        return 'str', None, b''.join(enc), off
    return None, None, None, None


def write_trec_arrays(path, qids, scores, rows, counts, docids, run_name, append=False, n_threads=0):
    """TREC run from [Q,k] arrays through the C++ writer (dhr_write_trec, csrc/trec.cu); rows index `docids`
    (shard-local), negative rows are padding.  Returns the number of lines written, or None when the id types are not
    plain ints / strs (the caller then formats in Python)."""
    qk, qi, qs, qo = _id_table(qids)
    dk, di, ds, do = _id_table(docids)
    if qk is None or dk is None:
        return None
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    nq, k = rows.shape
    cnt = None if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
    ptr = lambda a: None if a is None else a.ctypes.data
    n_lines = ctypes.c_int64(0)
    _cabi.check(_cabi.lib().dhr_write_trec(
        os.fsencode(path), 1 if append else 0, nq, k, ptr(cnt), ptr(rows), ptr(scores), ptr(qi), qs, ptr(qo),
        len(docids), ptr(di), ds, ptr(do), 1 if qk == dk else 0, str(run_name).encode('utf-8'), int(n_threads),
        ctypes.byref(n_lines)), 'dhr_write_trec')
    return n_lines.value


def write_trec(path, results, scores, docids, run_name):
    """gip_retrieval.py:329-342: rows whose docid equals the query id are skipped, ranks are NOT renumbered.
    The per-line formatting runs in C++ (same text, including Python's float repr); exotic id types fall back to the
    reference's Python loop."""
    qids = list(results.keys())
    kmax = max((len(results[q]) for q in qids), default=0)
    if qids and kmax > 0:
        rows = np.full((len(qids), kmax), -1, np.int64)
        sc = np.zeros((len(qids), kmax), np.float32)
        for i, q in enumerate(qids):
            n = len(results[q])
            rows[i, :n] = results[q]
            sc[i, :n] = scores[q]
        if write_trec_arrays(path, qids, sc, rows, None, docids, run_name) is not None:
            return
    with open(path, 'w') as fout:
        for query_id in results:
            result, score = results[query_id], scores[query_id]
            lines = ['{} Q0 {} {} {} {}\n'.format(query_id, docids[docidx], rank + 1, score[rank], run_name)
                     for rank, docidx in enumerate(result) if docids[docidx] != query_id]
            fout.write(''.join(lines))


def main(argv=None):
    args = build_parser().parse_args(argv)

    print('Load query embeddings ...')
    with open(args.query_emb_path, 'rb') as f:
        query_embs, query_arg_idxs, qids = pickle.load(f)
    query_embs = np.asarray(query_embs).astype(np.float32)          # :275
    if not isinstance(query_arg_idxs, np.ndarray):                   # :276-279: anything else means dense-only
        query_arg_idxs = None
    cls_dim = query_embs.shape[1] - args.emb_dim
    if cls_dim > 0:
        query_embs[:, -cls_dim:] = args.lamda * query_embs[:, -cls_dim:]   # :281-283

    print('Load index ...')
    if os.path.isdir(args.index_path):
        # mmap-able .npy container (dhr_b200/index_io.py): only this shard's rows are touched
        from .index_io import open_gip_index
        index, docids = open_gip_index(args.index_path, args.total_shrad, args.shrad, device=args.device,
                                       n_slices=args.emb_dim if query_arg_idxs is not None else 0, group=1)
    else:
        with open(args.index_path, 'rb') as f:
            corpus_embs, corpus_arg_idxs, docids = pickle.load(f)
        lo, hi = shard_bounds(len(docids), args.total_shrad, args.shrad)
        corpus_embs = corpus_embs[lo:hi]
        corpus_arg_idxs = corpus_arg_idxs[lo:hi] if isinstance(corpus_arg_idxs, np.ndarray) else None
        docids = docids[lo:hi]
        if query_arg_idxs is not None:
            index = GipIndex.from_arrays(corpus_embs, corpus_arg_idxs, n_slices=args.emb_dim, group=1, device=args.device)
        else:
            index = GipIndex.from_arrays(corpus_embs, None, device=args.device)
        del corpus_embs, corpus_arg_idxs

    if query_arg_idxs is not None and args.PQIP:                     # :321-322
        results, scores = PQ_IP_retrieval(qids, query_embs, query_arg_idxs, index, None, args)
    elif query_arg_idxs is not None:
        results, scores = GIP_retrieval(qids, query_embs, query_arg_idxs, index, None, args)
    else:
        results, scores = IP_retrieval(qids, query_embs, index, args)
    index.close()

    out = 'result.trec' if args.total_shrad == 1 else 'result{}.trec'.format(args.shrad)
    write_trec(out, results, scores, docids, args.run_name)
    print('finish')


if __name__ == "__main__":
    main()
