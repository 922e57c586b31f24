#!/usr/bin/env python
"""Executable statement (numpy, CPU) of the planned K1t v7 layout and walk -- "postings within the tile" (DESIGN.md section 7).

Not part of the product: a specification the CUDA kernel of the next round is written and tested against, plus the
occupancy / traffic model of the scheme on the benchmark's synthetic data.

Layout, per (tile of `tile_rows` passages, slice s):  the non-empty passage slices (some value != 0) sorted by code
(ties by passage id):  off[code] .. off[code + 1] delimit the items {passage id within the tile (u16), G fp16 values}.
Walk, per query tile:  for every slice s and query q with a non-empty slice, the matches of q in the tile are exactly the
items of list (s, code_q[s]);  acc[q][p] += sum_g qv[q,s,g] * pv[item,g]  (fp16 x fp16 products are exact in fp32).

    python tools/k1t_postings_spec.py            # model on 4 tiles of the delade_cls recipe
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_postings(c_lex, c_idx, S, G, rt, tile_rows=512):
    """c_lex [N, S*G] fp16 lexical values, c_idx [N, S] integer codes in [0, rt).
    Returns a list over tiles of dicts: off [S, rt + 1] int32, pid [n_items] uint16, val [n_items, G] fp16 (slice-major)."""
    N = c_lex.shape[0]
    v = np.asarray(c_lex).reshape(N, S, G)
    tiles = []
    for r0 in range(0, N, tile_rows):
        vt, it = v[r0:r0 + tile_rows], np.asarray(c_idx[r0:r0 + tile_rows]).astype(np.int64)
        off = np.zeros((S, rt + 1), np.int32)
        pids, vals = [], []
        base = 0
        for s in range(S):
            nz = np.any(vt[:, s, :] != 0, axis=1)                       # an all-zero slice never matches (0 * x = 0)
            p = np.nonzero(nz)[0]
            order = np.argsort(it[p, s], kind='stable')                  # by code, ties by passage id
            p = p[order]
            codes = it[p, s]
            cnt = np.bincount(codes, minlength=rt)[:rt]
            off[s, 0] = base
            off[s, 1:] = base + np.cumsum(cnt)
            base += len(p)
            pids.append(p.astype(np.uint16))
            vals.append(vt[p, s, :])
        tiles.append(dict(off=off, pid=np.concatenate(pids) if pids else np.zeros(0, np.uint16),
                          val=np.concatenate(vals) if vals else np.zeros((0, G), np.float16), rows=vt.shape[0]))
    return tiles


def walk_tile(tile, q_lex, q_idx, S, G, rt):
    """Lexical scores [Q, rows] (fp32) of one tile, accumulated slice by slice like the kernel; also the list lengths."""
    Q = q_lex.shape[0]
    qv = np.asarray(q_lex, dtype=np.float32).reshape(Q, S, G)
    qi = np.asarray(q_idx).astype(np.int64)
    acc = np.zeros((Q, tile['rows']), np.float32)
    lens = np.zeros((S, Q), np.int32)
    off, pid, val = tile['off'], tile['pid'], tile['val'].astype(np.float32)
    for s in range(S):
        for q in range(Q):
            c = qi[q, s]
            if c < 0 or c >= rt or not np.any(qv[q, s] != 0):             # empty query slice: contributes nothing
                continue
            a, b = off[s, c], off[s, c + 1]
            lens[s, q] = b - a
            if b > a:
                d = np.zeros(b - a, np.float32)
                for g in range(G):                                        # two FMA chains in the kernel; any fixed order is fine
                    d = d + qv[q, s, g] * val[a:b, g]
                acc[q, pid[a:b]] += d
    return acc, lens


def occupancy_model(lens, lists_per_warp=4):
    """Lane occupancy when a warp flattens the items of `lists_per_warp` lists (same slice, different queries) over its
    32 lanes: issued lane slots = 32 * ceil(items / 32) per group of lists."""
    S, Q = lens.shape
    used = issued = 0
    for s in range(S):
        nz = lens[s][lens[s] > 0]
        for i in range(0, len(nz), lists_per_warp):
            n = int(nz[i:i + lists_per_warp].sum())
            used += n
            issued += 32 * ((n + 31) // 32)
    return used / max(1, issued), used


def v6_occupancy(c_idx_tile, nonempty_tile, q_idx, q_nonempty, S, slices_per_chunk):
    """Lane occupancy of the current kernel (thread = passage, a warp steps to the largest per-thread match count of the
    chunk): matches / (32 * sum over (warp, chunk) of the warp maximum)."""
    rows = c_idx_tile.shape[0]
    used = issued = 0
    for c0 in range(0, S, slices_per_chunk):
        cnt = np.zeros(rows, np.int64)
        for s in range(c0, c0 + slices_per_chunk):
            m = (c_idx_tile[:, s][:, None] == q_idx[:, s][None, :]) & nonempty_tile[:, s][:, None] & q_nonempty[:, s][None, :]
            cnt += m.sum(axis=1)
        used += int(cnt.sum())
        for w in range(0, rows, 32):
            issued += 32 * int(cnt[w:w + 32].max())
    return used / max(1, issued)


def main():
    from dhr_b200 import synth
    from oracle import c_oracle
    cfg = synth.CONFIGS['delade_cls']
    S, G, C = cfg['S'], cfg['G'], cfg['C']
    rt = 39
    n_tiles, tile_rows, Q = 4, 512, 64
    cv, ci = synth.corpus_numpy('delade_cls', 0, n_tiles * tile_rows)
    qv, qi = synth.queries_numpy('delade_cls', Q)
    tiles = build_postings(cv[:, :S * G], ci, S, G, rt, tile_rows)
    items = sum(len(t['pid']) for t in tiles)
    print('layout: %.0f bytes per row (items %d B + offsets) vs %d B in the per-passage tile layout' % (
        (items * (2 + 2 * G) + n_tiles * S * (rt + 1) * 2) / (n_tiles * tile_rows), 2 + 2 * G, S * (1 + 2 * G)))
    worst = 0.0
    occ = []
    for t, tile in enumerate(tiles):
        acc, lens = walk_tile(tile, qv[:, :S * G], qi, S, G, rt)
        for q in range(0, Q, 16):                                         # check a few queries against the exact oracle
            z = np.zeros(C, np.float16)
            ex = c_oracle.scores(np.concatenate([cv[t * tile_rows:(t + 1) * tile_rows, :S * G], np.tile(z, (tile['rows'], 1))], axis=1),
                                 ci[t * tile_rows:(t + 1) * tile_rows], np.concatenate([qv[q, :S * G].astype(np.float32), z.astype(np.float32)]),
                                 qi[q], S, G)
            worst = max(worst, float(np.max(np.abs(acc[q].astype(np.float64) - ex))))
        for lw in (1, 2, 4, 8):
            occ.append((lw,) + occupancy_model(lens, lw))
        nz = lens[lens > 0]
        print('tile %d: %d matches, %d non-empty lists, mean list %.2f, max %d' % (t, int(lens.sum()), len(nz), nz.mean(), nz.max()))
    print('max |score - oracle| over the checked queries: %.2e' % worst)
    ne = np.any(cv[:, :S * G].reshape(-1, S, G) != 0, axis=2)
    qne = np.any(qv[:, :S * G].reshape(-1, S, G) != 0, axis=2)
    for sc in (4, 8, 16):
        o = np.mean([v6_occupancy(ci[t * tile_rows:(t + 1) * tile_rows].astype(np.int64), ne[t * tile_rows:(t + 1) * tile_rows],
                                  qi.astype(np.int64), qne, S, sc) for t in range(n_tiles)])
        print('current kernel (thread = passage), %2d slices per flattened walk: lane occupancy %.1f %%' % (sc, 100 * o))
    for lw in (1, 2, 4, 8):
        sel = [o for o in occ if o[0] == lw]
        print('lists per warp %d: lane occupancy %.1f %%' % (lw, 100 * np.mean([o[1] for o in sel])))


if __name__ == '__main__':
    main()
