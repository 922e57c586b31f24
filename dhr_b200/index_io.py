"""Index container next to the reference's pickle triple.

The reference stores an index as ``pickle.dump([values fp16 [N,W], idx [N,S] | 0, docids], protocol=4)``
(retrieval/index.py:46-47) and every shard process unpickles the WHOLE file before slicing
(gip_retrieval.py:289-306).  This module adds a lossless, mmap-able directory layout

    <dir>/values.npy   fp16 [N, W]          <dir>/idx.npy   integer [N, S]   (absent for dense-only indexes)
    <dir>/docids.txt   one id per line      <dir>/meta.json {"n": N, "width": W, "n_slices": S, "group": G}

so a rank maps only its own row range and streams it to the GPU in chunks (no rank ever holds the
whole index in host memory).  ``python -m dhr_b200.index_io to-npy index.pt out_dir`` / ``to-pickle`` convert
both ways; the merge of encoder splits (retrieval/index.py:26-47) is ``merge-splits``.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import pickle

import numpy as np

from .gip_retrieval import shard_bounds


def pickle_to_npy(index_path, out_dir, n_slices=None, group=1):
    with open(index_path, 'rb') as f:
        values, idx, docids = pickle.load(f)
    save_npy(out_dir, values, idx if isinstance(idx, np.ndarray) else None, docids, n_slices, group)


def save_npy(out_dir, values, idx, docids, n_slices=None, group=1):
    os.makedirs(out_dir, exist_ok=True)
    values = np.ascontiguousarray(values)
    np.save(os.path.join(out_dir, 'values.npy'), values)
    S = 0
    if idx is not None:
        idx = np.ascontiguousarray(idx)
        S = idx.shape[1] if n_slices is None else n_slices
        np.save(os.path.join(out_dir, 'idx.npy'), idx)
    with open(os.path.join(out_dir, 'docids.txt'), 'w') as f:
        f.write('\n'.join(str(d) for d in docids))
        f.write('\n' if len(docids) else '')
    with open(os.path.join(out_dir, 'meta.json'), 'w') as f:
        json.dump({'n': int(values.shape[0]), 'width': int(values.shape[1]), 'n_slices': int(S), 'group': int(group),
                   'docid_type': 'int' if all(isinstance(d, (int, np.integer)) for d in docids) else 'str'}, f)


def npy_to_pickle(in_dir, index_path):
    values, idx, docids, _ = load_npy(in_dir, mmap=False)
    with open(index_path, 'wb') as f:
        pickle.dump([np.asarray(values), np.asarray(idx) if idx is not None else 0, docids], f, protocol=4)


def load_npy(in_dir, total_shrad=1, shrad=0, mmap=True):
    """(values, idx | None, docids, lo) of shard `shrad` under the reference's rule; arrays are memory-mapped views."""
    with open(os.path.join(in_dir, 'meta.json')) as f:
        meta = json.load(f)
    lo, hi = shard_bounds(meta['n'], total_shrad, shrad)
    values = np.load(os.path.join(in_dir, 'values.npy'), mmap_mode='r' if mmap else None)[lo:hi]
    ip = os.path.join(in_dir, 'idx.npy')
    idx = np.load(ip, mmap_mode='r' if mmap else None)[lo:hi] if os.path.exists(ip) else None
    with open(os.path.join(in_dir, 'docids.txt')) as f:
        docids = f.read().split('\n')
    if docids and docids[-1] == '':
        docids.pop()
    if meta.get('docid_type') == 'int':
        docids = [int(d) for d in docids]
    return values, idx, docids[lo:hi], lo


def open_gip_index(in_dir, total_shrad=1, shrad=0, device=0, n_slices=None, group=None, chunk_rows=1 << 18):
    """Build the HBM-resident shard straight from the memory-mapped files, `chunk_rows` rows at a time."""
    from .index import GipIndex
    values, idx, docids, lo = load_npy(in_dir, total_shrad, shrad)
    with open(os.path.join(in_dir, 'meta.json')) as f:
        meta = json.load(f)
    S = meta['n_slices'] if n_slices is None else n_slices
    G = meta['group'] if group is None else group
    if idx is None:
        S = 0
    ix = GipIndex(S, values.shape[1] - S * G, G, capacity=values.shape[0], idx_dtype=idx.dtype if idx is not None else np.uint8,
                  device=device, row_offset=lo)
    for r0 in range(0, values.shape[0], chunk_rows):
        r1 = min(values.shape[0], r0 + chunk_rows)
        ix.append(np.ascontiguousarray(values[r0:r1]), np.ascontiguousarray(idx[r0:r1]) if idx is not None else None)
    ix.finalize()
    return ix, docids


def merge_splits(index_path, index_prefix='msmarco-passage', out=None):
    """retrieval/index.py:26-47 with a deterministic (sorted) split order; writes <prefix>.index.pt."""
    files = sorted(glob.glob(os.path.join(index_path, index_prefix + '.split*.pt')))
    embs, idxs, docids = [], [], []
    for fn in files:
        with open(fn, 'rb') as f:
            e, i, d = pickle.load(f)
        embs.append(e)
        idxs.append(i)
        docids += list(d)
    try:
        idx = np.concatenate(idxs, axis=0)
    except Exception:                      # dense / agg models store None: the reference writes the int 0 (index.py:40-43)
        idx = 0
    out = out or os.path.join(index_path, index_prefix + '.index.pt')
    with open(out, 'wb') as f:
        pickle.dump([np.concatenate(embs, axis=0), idx, docids], f, protocol=4)
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    sub = ap.add_subparsers(dest='cmd', required=True)
    a = sub.add_parser('to-npy'); a.add_argument('index_pt'); a.add_argument('out_dir')
    a.add_argument('--emb_dim', type=int, default=None); a.add_argument('--group', type=int, default=1)
    b = sub.add_parser('to-pickle'); b.add_argument('in_dir'); b.add_argument('index_pt')
    c = sub.add_parser('merge-splits'); c.add_argument('--index_path', required=True)
    c.add_argument('--index_prefix', default='msmarco-passage'); c.add_argument('--emb_dim', type=int, default=768)
    args = ap.parse_args(argv)
    if args.cmd == 'to-npy':
        pickle_to_npy(args.index_pt, args.out_dir, args.emb_dim, args.group)
    elif args.cmd == 'to-pickle':
        npy_to_pickle(args.in_dir, args.index_pt)
    else:
        print(merge_splits(args.index_path, args.index_prefix))


if __name__ == '__main__':
    main()
