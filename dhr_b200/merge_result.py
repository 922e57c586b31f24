"""Mirror of castorini/dhr ``retrieval/merge.result.py`` (:14-43): merge per-shard TREC files into result.trec.

Same flags (--total_shrad, --topk, --run_name).  The reference reads ``result{:02d}.trec`` although
``gip_retrieval.py:332`` writes ``result{}.trec``; both spellings are accepted here.  Parsing, the per-query selection
and the formatting run on the host in C++ (``dhr_merge_trec``, csrc/trec.cu); ties are ordered by (score desc, position
in the concatenated shard lists asc), i.e. by shard then by the shard's own rank -- which equals the single-shard order
because every shard lists ties by row (the reference's ``argsort()[::-1]`` leaves the order of equal scores to numpy's
unstable sort)."""
from __future__ import annotations

import argparse
import ctypes
import os

from . import _cabi


def shard_paths(total_shrad, directory='.'):
    paths = []
    for shrad in range(total_shrad):
        for name in ('result{:02d}.trec'.format(shrad), 'result{}.trec'.format(shrad)):
            path = os.path.join(directory, name)
            if os.path.exists(path):
                paths.append(path)
                break
        else:
            raise FileNotFoundError('no result file for shard %d in %s' % (shrad, directory))
    return paths


def merge_trec(paths, out_path, topk, run_name, n_threads=0):
    """merge the TREC files `paths` (shard order) into `out_path`; returns the number of lines written"""
    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    n_lines = ctypes.c_int64(0)
    _cabi.check(_cabi.lib().dhr_merge_trec(len(paths), arr, os.fsencode(out_path), int(topk), str(run_name).encode('utf-8'),
                                           int(n_threads), ctypes.byref(n_lines)), 'dhr_merge_trec')
    return n_lines.value


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--total_shrad", type=int, default=1)
    parser.add_argument("--topk", type=int, default=1000)
    parser.add_argument("--run_name", default='dhr')
    args = parser.parse_args(argv)
    paths = shard_paths(args.total_shrad)
    print('write results ...')
    merge_trec(paths, 'result.trec', args.topk, args.run_name)


if __name__ == "__main__":
    main()
