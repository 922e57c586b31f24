#!/usr/bin/env python
"""Kernel-level timing loop for the tile-path scan kernels (K1t alone on a lexical-only shape, or K2 + K1t on a hybrid one):
a small resident index, one super-batch of queries, CUDA-event scan time per launch, result checked against the row scan K1.

    python tools/k1t_bench.py [--workload delade_lex] [--rows 606208] [--queries 256] [--reps 5] [--option name=value ...]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import numpy as np
import torch

from dhr_b200 import GipIndex, synth

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='delade_lex')
ap.add_argument('--rows', type=int, default=16 * 37888)
ap.add_argument('--queries', type=int, default=256)
ap.add_argument('--topk', type=int, default=1000)
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--option', action='append', default=[])
ap.add_argument('--no-check', action='store_true')
ap.add_argument('--lex-postings', action='store_true', help='experimental postings lexical layout (K1p) instead of the tiled one (K1t)')
a = ap.parse_args()

cfg = synth.CONFIGS[a.workload]
dev = torch.device('cuda', 0)
ix = GipIndex(cfg['S'], cfg['C'], cfg['G'], capacity=a.rows, idx_dtype=np.dtype(cfg['idx']), device=0, lex_postings=a.lex_postings)
for v, i in synth.corpus_torch_segments(a.workload, 0, a.rows, dev):
    ix.append(v, i)
ix.finalize()
for kv in a.option:
    n, v = kv.split('=')
    ix.set_option(n, int(v))
ix.set_option('profile', 1)
qv, qi = synth.queries_torch(a.workload, a.queries, dev)
out = None
best = None
for r in range(a.reps + 2):
    out = ix.search(qv, qi, a.topk)
    st = ix.stats()
    if r >= 2 and (best is None or st['scan_ms'] < best['scan_ms']):
        best = st
res = {'workload': a.workload, 'rows': a.rows, 'queries': a.queries, 'scan_ms': best['scan_ms'], 'select_ms': best['select_ms'],
       'total_ms': best['total_ms'], 'scan_launches': best['n_scan_launches'], 'us_per_scan_launch': 1e3 * best['scan_ms'] / max(1, best['n_scan_launches']),
       'scan_variant': best['scan_variant'], 'lex_layout': best['lex_layout'], 'q_per_s_at_8p8M': a.queries / (best['total_ms'] / 1e3) * a.rows / synth.N_MSMARCO}
if not a.no_check:
    n = min(16, a.queries)
    ix.set_option('tile_mode', 0)
    ref = ix.search(qv[:n], qi[:n] if qi is not None else None, a.topk)
    res['rows_equal_frac_vs_K1'] = float((ref[1] == out[1][:n]).float().mean().item())
    res['max_abs_score_diff_vs_K1'] = float((ref[0] - out[0][:n]).abs().max().item())
print(json.dumps(res))
