"""ctypes loader for oracle/_build/libgip_oracle.so.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libgip_oracle.so')
_IDX = {np.dtype('uint8'): 1, np.dtype('int8'): 2, np.dtype('int16'): 3, np.dtype('uint16'): 4,
        np.dtype('int32'): 5, np.dtype('int64'): 6}
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'gip_oracle.c')):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.gip_oracle_search.restype = ctypes.c_int
        _lib.gip_oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().gip_oracle_num_threads())


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def search(c_vals_f16, c_idx, q_vals_f32, q_idx, n_slices, group, k, masked=True):
    """Exact (fp64) top-k via the C oracle.  Returns (rows int64 [Q,k], scores f64 [Q,k])."""
    cv = np.ascontiguousarray(c_vals_f16, dtype=np.float16)
    qv = np.ascontiguousarray(q_vals_f32, dtype=np.float32)
    N, W = cv.shape
    Q = qv.shape[0]
    C = W - n_slices * group
    ci = np.ascontiguousarray(c_idx) if (c_idx is not None and n_slices > 0) else None
    qi = np.ascontiguousarray(q_idx) if (q_idx is not None and n_slices > 0) else None
    out_s = np.empty((Q, k), dtype=np.float64)
    out_r = np.empty((Q, k), dtype=np.int64)
    rc = lib().gip_oracle_search(
        ctypes.c_int64(N), ctypes.c_int(n_slices), ctypes.c_int(group), ctypes.c_int(C),
        _ptr(cv.view(np.uint16)), _ptr(ci), ctypes.c_int(_IDX[ci.dtype] if ci is not None else 0),
        ctypes.c_int(Q), _ptr(qv), _ptr(qi), ctypes.c_int(_IDX[qi.dtype] if qi is not None else 0),
        ctypes.c_int(1 if masked else 0), ctypes.c_int(k), _ptr(out_s), _ptr(out_r))
    if rc != 0:
        raise MemoryError('gip_oracle_search failed')
    return out_r, out_s


def scores(c_vals_f16, c_idx, q_val_row_f32, q_idx_row, n_slices, group, masked=True):
    """Exact fp64 scores [N] of one query."""
    cv = np.ascontiguousarray(c_vals_f16, dtype=np.float16)
    qv = np.ascontiguousarray(q_val_row_f32, dtype=np.float32)
    N, W = cv.shape
    C = W - n_slices * group
    ci = np.ascontiguousarray(c_idx) if (c_idx is not None and n_slices > 0) else None
    qi = np.ascontiguousarray(q_idx_row) if (q_idx_row is not None and n_slices > 0) else None
    out = np.empty(N, dtype=np.float64)
    lib().gip_oracle_scores(
        ctypes.c_int64(N), ctypes.c_int(n_slices), ctypes.c_int(group), ctypes.c_int(C),
        _ptr(cv.view(np.uint16)), _ptr(ci), ctypes.c_int(_IDX[ci.dtype] if ci is not None else 0),
        _ptr(qv), _ptr(qi), ctypes.c_int(_IDX[qi.dtype] if qi is not None else 0), ctypes.c_int64(0),
        ctypes.c_int(1 if masked else 0), _ptr(out))
    return out
