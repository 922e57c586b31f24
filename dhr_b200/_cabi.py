"""ctypes binding of include/dhr_b200.h (libdhr_b200.so).

There is no CPU fallback: importing this module without the built CUDA library raises
ImportError, and every call that fails raises DhrError with the C-ABI status text.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libdhr_b200.so')

# status codes / enums (mirror of include/dhr_b200.h)
OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_LOSSY, ERR_IDX_RANGE, ERR_STATE, ERR_NO_DEVICE = range(1, 9)
IDX_NONE, IDX_U8, IDX_I8, IDX_I16, IDX_U16, IDX_I32, IDX_I64 = range(7)
VAL_F16, VAL_F32 = 0, 1
SEARCH_UNMASKED = 1
INDEX_NARROW_CODES = 1
INDEX_KEEP_ROWMAJOR = 2
INDEX_LEX_POSTINGS = 4
MAX_K = 12288
MAX_GROUP = 8

EXPORTS = [
    'dhr_version', 'dhr_strerror', 'dhr_last_cuda_error', 'dhr_device_count',
    'dhr_index_create', 'dhr_index_append', 'dhr_index_finalize', 'dhr_index_open', 'dhr_index_close',
    'dhr_index_rows', 'dhr_index_row_bytes', 'dhr_index_set_option', 'dhr_index_get_stats',
    'dhr_search', 'dhr_rerank', 'dhr_topk_merge', 'dhr_densify', 'dhr_write_trec', 'dhr_merge_trec',
    'dhr_search_keys', 'dhr_search_batches', 'dhr_search_wait_batch', 'dhr_search_complete', 'dhr_merge_keys',
    'dhr_index_device_bytes',
]


class DhrStats(ctypes.Structure):
    _fields_ = [
        ('n_queries', ctypes.c_int32), ('query_block', ctypes.c_int32), ('query_groups', ctypes.c_int32),
        ('scan_variant', ctypes.c_int32), ('n_scan_launches', ctypes.c_int32), ('n_select_launches', ctypes.c_int32),
        ('n_prep_launches', ctypes.c_int32), ('n_fallback_queries', ctypes.c_int32),
        ('n_kernel_launches', ctypes.c_int32), ('rowmajor_rebuilds', ctypes.c_int32),
        ('scan_ms', ctypes.c_double), ('select_ms', ctypes.c_double), ('total_ms', ctypes.c_double),
        ('corpus_passes', ctypes.c_double), ('bytes_per_pass', ctypes.c_double), ('dense_flops', ctypes.c_double), ('lex_layout', ctypes.c_double), ('alg_bytes', ctypes.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class DhrError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib().dhr_strerror(status).decode()
        if status == ERR_CUDA:
            msg += ': ' + lib().dhr_last_cuda_error().decode()
        super().__init__('%s failed: %s (status %d)' % (where, msg, status))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'dhr_b200: CUDA library %s is missing. Build it with `python -m dhr_b200.build` '
            '(needs nvcc; there is no CPU fallback).' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i64, i32, f32 = c.c_void_p, c.c_int64, c.c_int, c.c_float
    L.dhr_version.restype = i32
    L.dhr_strerror.restype = c.c_char_p
    L.dhr_strerror.argtypes = [i32]
    L.dhr_last_cuda_error.restype = c.c_char_p
    L.dhr_device_count.argtypes = [c.POINTER(i32)]
    L.dhr_index_create.argtypes = [c.POINTER(vp), i32, i64, i32, i32, i32, i32, i64, c.c_uint]
    L.dhr_index_append.argtypes = [vp, i64, i32, vp, i64, i32, vp, i64]
    L.dhr_index_finalize.argtypes = [vp]
    L.dhr_index_open.argtypes = [c.POINTER(vp), i32, i64, i32, i32, i32, i32, vp, i64, i32, vp, i64, i64, c.c_uint]
    L.dhr_index_close.argtypes = [vp]
    L.dhr_index_rows.argtypes = [vp, c.POINTER(i64)]
    L.dhr_index_row_bytes.argtypes = [vp, c.POINTER(i64)]
    L.dhr_index_set_option.argtypes = [vp, c.c_char_p, i64]
    L.dhr_index_get_stats.argtypes = [vp, c.POINTER(DhrStats)]
    L.dhr_search.argtypes = [vp, i32, i32, vp, i64, i32, vp, i64, f32, i32, c.c_uint, vp, vp, vp, vp]
    L.dhr_search_keys.argtypes = [vp, i32, i32, vp, i64, i32, vp, i64, f32, i32, c.c_uint, vp, vp]
    L.dhr_search_batches.argtypes = [vp, c.POINTER(i32), c.POINTER(i32)]
    L.dhr_search_wait_batch.argtypes = [vp, i32, vp]
    L.dhr_search_complete.argtypes = [vp, c.POINTER(i32), vp]
    L.dhr_merge_keys.argtypes = [i32, i32, i32, i32, vp, i64, vp, vp, vp, vp]
    L.dhr_index_device_bytes.argtypes = [vp, c.POINTER(i64)]
    L.dhr_rerank.argtypes = [vp, i32, i32, vp, i64, i32, vp, i64, f32, vp, i32, i32, vp, vp, vp, vp]
    L.dhr_topk_merge.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.dhr_densify.argtypes = [i32, i32, i32, i32, i32, i32, vp, i64, vp, i64, vp, i64, vp]
    L.dhr_write_trec.argtypes = [c.c_char_p, i32, i32, i32, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, i32, c.c_char_p, i32, c.POINTER(i64)]
    L.dhr_merge_trec.argtypes = [i32, c.POINTER(c.c_char_p), c.c_char_p, i32, c.c_char_p, i32, c.POINTER(i64)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is c.c_int and name not in ('dhr_version',):
            fn.restype = i32
    _lib = L
    return L


def check(status, where):
    if status != OK:
        raise DhrError(status, where)


def device_count():
    n = ctypes.c_int(0)
    st = lib().dhr_device_count(ctypes.byref(n))
    return n.value if st == OK else 0
