"""Synthetic encoded corpora / queries of MS MARCO shape, for bench.py and the tests.

Rows are generated in fixed segments of SEG rows, segment j from seed (seed + j), so a range shard
(any rank count) sees exactly the rows a single-GPU run sees.  numpy on the host for tests, torch on
the device for the full-size bench (the 29 GB corpus never exists on the host).

Recipes (per config of BASELINE.json):
  delade  lexical fp16 |N(0, 0.2)|, 30 % of slices empty (value 0, idx 0), idx uniform in [0, 39)
          (DHR/utils.py:20-21: 29952 / 768 = 39), dense fp16 N(0, 1)/sqrt(C)
  bm25    5 % of slices non-empty, value fp16 U(0.5, 8), idx uniform in [0, 3466); queries have
          <= 8 non-empty slices with small-integer tf values (densify_query.py:87-89) -> exact ties
  dense   fp16 N(0, 1)/sqrt(C), no lexical part (Aggretriever)
"""
from __future__ import annotations

import numpy as np

SEG = 1 << 16

CONFIGS = {
    # name: (S, G, C, idx dtype, recipe, idx range)
    'delade_cls': dict(S=128, G=6, C=768, idx='uint16', recipe='delade', R=39),      # BASELINE config 2 (literal)
    'delade_cls_ref': dict(S=768, G=1, C=128, idx='uint8', recipe='delade', R=39),   # reference-true shape (2')
    'bm25': dict(S=256, G=3, C=0, idx='uint16', recipe='bm25', R=3466),              # config 3 (literal)
    'bm25_ref': dict(S=768, G=1, C=0, idx='int16', recipe='bm25', R=3466),           # config 3'
    'dense': dict(S=0, G=1, C=768, idx='uint8', recipe='dense', R=1),                # config 4
    # robustness variant of config 2: slice indices Zipf-distributed over the 39 values (p(r) ~ 1/(r+1)), like the argmax over
    # vocabulary strides of a real encoder where a few strides win most often -> ~3.5x the matches of the uniform recipe
    'delade_cls_zipf': dict(S=128, G=6, C=768, idx='uint16', recipe='delade', R=39, zipf=1.0),
    # kernel-level variants (tools/k1t_bench.py): the lexical part of configs 2 / 2' alone
    'delade_lex': dict(S=128, G=6, C=0, idx='uint16', recipe='delade', R=39),
    'delade_ref_lex': dict(S=768, G=1, C=0, idx='uint8', recipe='delade', R=39),
}
N_MSMARCO = 8841823
Q_MSMARCO = 6808


def _segments(lo, hi):
    j = lo // SEG
    while j * SEG < hi:
        a, b = max(lo, j * SEG), min(hi, (j + 1) * SEG)
        yield j, a - j * SEG, b - j * SEG
        j += 1


# ---- numpy (host) ------------------------------------------------------------------------------------
def _np_segment(cfg, seed, n, queries):
    rng = np.random.default_rng(seed)
    S, G, C, R = cfg['S'], cfg['G'], cfg['C'], cfg['R']
    parts = []
    idx = np.zeros((n, S), dtype=cfg['idx'])
    if S > 0:
        if cfg['recipe'] == 'delade':
            v = np.abs(rng.standard_normal((n, S, G), dtype=np.float32) * 0.2)
            empty = rng.random((n, S)) < 0.3
        else:
            if queries:
                v = np.floor(rng.random((n, S, G), dtype=np.float32) * 3 + 1)           # tf in {1,2,3}
                empty = np.ones((n, S), bool)
                for r in range(n):
                    empty[r, rng.choice(S, size=min(S, 8), replace=False)] = False
            else:
                v = rng.random((n, S, G), dtype=np.float32) * 7.5 + 0.5
                empty = rng.random((n, S)) >= 0.05
        if cfg.get('zipf'):
            pr = 1.0 / np.arange(1, R + 1, dtype=np.float64) ** cfg['zipf']
            ii = rng.choice(R, size=(n, S), p=pr / pr.sum())
        else:
            ii = rng.integers(0, R, size=(n, S))
        v[empty] = 0
        ii[empty] = 0
        parts.append(v.reshape(n, S * G).astype(np.float16))
        idx = ii.astype(cfg['idx'])
    if C > 0:
        parts.append((rng.standard_normal((n, C), dtype=np.float32) / np.sqrt(C)).astype(np.float16))
    return np.concatenate(parts, axis=1), idx


def corpus_numpy(cfg, lo, hi, seed=1234):
    """Rows [lo, hi) of the synthetic corpus: (vals fp16 [n, W], idx [n, S])."""
    cfg = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    vs, is_ = [], []
    for j, a, b in _segments(lo, hi):
        v, i = _np_segment(cfg, seed + j, SEG, False)
        vs.append(v[a:b])
        is_.append(i[a:b])
    if not vs:
        W = cfg['S'] * cfg['G'] + cfg['C']
        return np.zeros((0, W), np.float16), np.zeros((0, cfg['S']), cfg['idx'])
    return np.concatenate(vs), np.concatenate(is_)


def queries_numpy(cfg, n, seed=4321):
    cfg = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return _np_segment(cfg, seed, n, True)


# ---- torch (device) ----------------------------------------------------------------------------------
def _torch_idx_dtype(torch, name):
    return {'uint8': torch.uint8, 'int8': torch.int8, 'int16': torch.int16, 'uint16': torch.uint16}[name]


def _torch_segment(torch, cfg, seed, n, queries, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    S, G, C, R = cfg['S'], cfg['G'], cfg['C'], cfg['R']
    parts, idx = [], None
    if S > 0:
        if cfg['recipe'] == 'delade':
            v = (torch.randn((n, S, G), generator=g, device=device, dtype=torch.float32) * 0.2).abs_()
            empty = torch.rand((n, S), generator=g, device=device) < 0.3
        else:
            if queries:
                v = torch.floor(torch.rand((n, S, G), generator=g, device=device) * 3 + 1)
                keep = torch.rand((n, S), generator=g, device=device).argsort(dim=1)[:, :min(S, 8)]
                empty = torch.ones((n, S), dtype=torch.bool, device=device)
                empty.scatter_(1, keep, False)
            else:
                v = torch.rand((n, S, G), generator=g, device=device) * 7.5 + 0.5
                empty = torch.rand((n, S), generator=g, device=device) >= 0.05
        if cfg.get('zipf'):
            pr = 1.0 / torch.arange(1, R + 1, device=device, dtype=torch.float32) ** cfg['zipf']
            ii = torch.multinomial(pr / pr.sum(), n * S, replacement=True, generator=g).reshape(n, S).to(torch.int32)
        else:
            ii = torch.randint(0, R, (n, S), generator=g, device=device, dtype=torch.int32)
        v[empty] = 0
        ii[empty] = 0
        parts.append(v.reshape(n, S * G).to(torch.float16))
        idx = ii.to(torch.int16).view(_torch_idx_dtype(torch, cfg['idx'])) if cfg['idx'] in ('uint16', 'int16') \
            else ii.to(_torch_idx_dtype(torch, cfg['idx']))
    if C > 0:
        parts.append((torch.randn((n, C), generator=g, device=device, dtype=torch.float32) / (C ** 0.5)).to(torch.float16))
    vals = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]
    return vals, idx


def corpus_torch_segments(cfg, lo, hi, device, seed=1234):
    """Yield (vals, idx) CUDA tensors covering rows [lo, hi) segment by segment (bounded memory)."""
    import torch
    cfg = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    for j, a, b in _segments(lo, hi):
        v, i = _torch_segment(torch, cfg, seed + j, SEG, False, device)
        yield v[a:b], (i[a:b] if i is not None else None)


def queries_torch(cfg, n, device, seed=4321):
    import torch
    cfg = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return _torch_segment(torch, cfg, seed, n, True, device)
