"""Mirror of castorini/dhr ``retrieval/merge.result.py`` (:14-43): merge per-shard TREC files into result.trec.

Same flags (--total_shrad, --topk, --run_name).  The reference reads ``result{:02d}.trec`` although
``gip_retrieval.py:332`` writes ``result{}.trec``; both spellings are accepted here.  The per-query selection runs on
the GPU through ``dhr_topk_merge`` and orders ties by (score desc, position in the concatenated shard lists asc), i.e.
by shard then by the shard's own rank -- which equals the single-shard order because every shard lists ties by row."""
from __future__ import annotations

import argparse
import os
from collections import OrderedDict

import numpy as np

from .index import topk_merge


def read_shards(total_shrad, directory='.'):
    per_q = OrderedDict()
    for shrad in range(total_shrad):
        for name in ('result{:02d}.trec'.format(shrad), 'result{}.trec'.format(shrad)):
            path = os.path.join(directory, name)
            if os.path.exists(path):
                break
        else:
            raise FileNotFoundError('no result file for shard %d in %s' % (shrad, directory))
        with open(path) as f:
            for line in f:
                query_id, _, docid, _rank, score, _ = line.strip().split(' ')
                per_q.setdefault(query_id, []).append((shrad, docid, score))
    return per_q


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--total_shrad", type=int, default=1)
    parser.add_argument("--topk", type=int, default=1000)
    parser.add_argument("--run_name", default='dhr')
    parser.add_argument("--device", type=int, default=0)
    args = parser.parse_args(argv)
    per_q = read_shards(args.total_shrad)
    qids = list(per_q.keys())
    width = max((len(v) for v in per_q.values()), default=0)
    if width == 0:
        open('result.trec', 'w').close()
        return
    # one "part" holding, per query, all shards' candidates in file order; position encodes the tie order
    scores = np.full((1, len(qids), width), -np.inf, np.float32)
    rows = np.full((1, len(qids), width), -1, np.int64)
    for i, q in enumerate(qids):
        n = len(per_q[q])
        scores[0, i, :n] = [float(s) for _, _, s in per_q[q]]
        rows[0, i, :n] = np.arange(n)
    print('write results ...')
    ms, mr = topk_merge(scores, rows, device=args.device)
    with open('result.trec', 'w') as fout:
        for i, q in enumerate(qids):
            out = []
            for rank in range(min(args.topk, width)):
                pos = int(mr[i, rank])
                if pos < 0:
                    break
                _, docid, score_text = per_q[q][pos]
                out.append('{} Q0 {} {} {} {}\n'.format(q, docid, rank + 1, float(score_text), args.run_name))
            fout.write(''.join(out))


if __name__ == "__main__":
    main()
