"""GipIndex: HBM-resident GIP index + search, the Python face of the C ABI (include/dhr_b200.h).

Arrays may be numpy arrays or torch tensors (CPU or CUDA); PyTorch is only the memory carrier,
the arithmetic is in libdhr_b200.so.  Replaces the load / H2D / scoring / top-k of
castorini/dhr retrieval/gip_retrieval.py:289-327.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi as C

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

_NP_IDX = {np.dtype('uint8'): C.IDX_U8, np.dtype('int8'): C.IDX_I8, np.dtype('int16'): C.IDX_I16,
           np.dtype('uint16'): C.IDX_U16, np.dtype('int32'): C.IDX_I32, np.dtype('int64'): C.IDX_I64}
_NP_VAL = {np.dtype('float16'): C.VAL_F16, np.dtype('float32'): C.VAL_F32}


def _torch_maps():
    idx = {torch.uint8: C.IDX_U8, torch.int8: C.IDX_I8, torch.int16: C.IDX_I16, torch.int32: C.IDX_I32,
           torch.int64: C.IDX_I64}
    if hasattr(torch, 'uint16'):
        idx[torch.uint16] = C.IDX_U16
    val = {torch.float16: C.VAL_F16, torch.float32: C.VAL_F32}
    return idx, val


class _Arr:
    """pointer / dtype code / row stride (elements) of a 2-D array whose rows are contiguous"""

    def __init__(self, a, kind):
        self.keep = a
        if torch is not None and isinstance(a, torch.Tensor):
            idx_map, val_map = _torch_maps()
            table = idx_map if kind == 'idx' else val_map
            if a.dtype not in table:
                raise TypeError('unsupported %s dtype %s' % (kind, a.dtype))
            if a.dim() != 2:
                raise ValueError('expected a 2-D array, got shape %s' % (tuple(a.shape),))
            if a.shape[1] > 1 and a.stride(1) != 1:
                a = a.contiguous()
                self.keep = a
            self.ptr, self.code = a.data_ptr(), table[a.dtype]
            self.stride = a.stride(0) if a.shape[0] > 1 else max(a.shape[1], 1)
            self.shape = tuple(a.shape)
            self.device = a.device
        else:
            a = np.asarray(a)
            table = _NP_IDX if kind == 'idx' else _NP_VAL
            if a.dtype not in table:
                raise TypeError('unsupported %s dtype %s' % (kind, a.dtype))
            if a.ndim != 2:
                raise ValueError('expected a 2-D array, got shape %s' % (a.shape,))
            if a.shape[1] > 1 and a.strides[1] != a.itemsize or a.strides[0] % a.itemsize or a.strides[0] < 0:
                a = np.ascontiguousarray(a)
            self.keep = a
            self.ptr, self.code = a.ctypes.data, table[a.dtype]
            self.stride = a.strides[0] // a.itemsize if a.shape[0] > 1 else max(a.shape[1], 1)
            self.shape = a.shape
            self.device = None
        if self.stride < self.shape[1]:
            self.stride = self.shape[1]


class GipIndex:
    """One corpus shard resident in the HBM of one GPU."""

    def __init__(self, n_slices, n_dense, group=1, capacity=0, idx_dtype=np.uint8, device=0, row_offset=0,
                 narrow_codes=False, keep_rowmajor=False, lex_postings=False):
        self._h = ctypes.c_void_p()
        self.n_slices, self.n_dense, self.group = int(n_slices), int(n_dense), int(group)
        self.device, self.row_offset = int(device), int(row_offset)
        self.width = self.n_slices * self.group + self.n_dense
        code = C.IDX_NONE if n_slices == 0 else _NP_IDX[np.dtype(idx_dtype)]
        flags = (C.INDEX_NARROW_CODES if narrow_codes else 0) | (C.INDEX_KEEP_ROWMAJOR if keep_rowmajor else 0) | (C.INDEX_LEX_POSTINGS if lex_postings else 0)
        C.check(C.lib().dhr_index_create(ctypes.byref(self._h), self.device, int(capacity), self.n_slices, self.group,
                                         self.n_dense, code, self.row_offset, flags), 'dhr_index_create')

    # ---- construction --------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, vals, idx, n_slices=None, group=1, device=0, row_offset=0, narrow_codes=False, keep_rowmajor=False, lex_postings=False):
        """vals [N, S*G + C] fp16/fp32, idx [N, S] integer array or None / 0 (dense-only index,
        the reference stores None or 0 there: encode.py:149-153, index.py:40-43)."""
        v = _Arr(vals, 'val')
        has_idx = idx is not None and not np.isscalar(idx) and getattr(idx, 'ndim', 2) == 2 and idx.shape[1] > 0
        if has_idx:
            i = _Arr(idx, 'idx')
            S = i.shape[1] if n_slices is None else int(n_slices)
            np_dtype = {c: d for d, c in _NP_IDX.items()}[i.code]
        else:
            S, np_dtype = 0, np.uint8
        C_ = v.shape[1] - S * group
        if C_ < 0:
            raise ValueError('values have %d columns but n_slices*group = %d' % (v.shape[1], S * group))
        self = cls(S, C_, group, capacity=v.shape[0], idx_dtype=np_dtype, device=device, row_offset=row_offset,
                   narrow_codes=narrow_codes, keep_rowmajor=keep_rowmajor, lex_postings=lex_postings)
        self.append(vals, idx if has_idx else None)
        self.finalize()
        return self

    def append(self, vals, idx=None):
        v = _Arr(vals, 'val')
        if v.shape[1] != self.width:
            raise ValueError('values have %d columns, index expects %d' % (v.shape[1], self.width))
        if self.n_slices > 0:
            if idx is None:
                raise ValueError('index has %d slices: idx is required' % self.n_slices)
            i = _Arr(idx, 'idx')
            if i.shape != (v.shape[0], self.n_slices):
                raise ValueError('idx shape %s, expected %s' % (i.shape, (v.shape[0], self.n_slices)))
            ip, ic, istr = i.ptr, i.code, i.stride
        else:
            ip, ic, istr = None, C.IDX_NONE, 0
        C.check(C.lib().dhr_index_append(self._h, v.shape[0], v.code, v.ptr, v.stride, ic, ip, istr), 'dhr_index_append')

    def finalize(self):
        C.check(C.lib().dhr_index_finalize(self._h), 'dhr_index_finalize')

    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            C.lib().dhr_index_close(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- properties ------------------------------------------------------------------------------
    def __len__(self):
        n = ctypes.c_int64(0)
        C.check(C.lib().dhr_index_rows(self._h, ctypes.byref(n)), 'dhr_index_rows')
        return n.value

    @property
    def row_bytes(self):
        n = ctypes.c_int64(0)
        C.check(C.lib().dhr_index_row_bytes(self._h, ctypes.byref(n)), 'dhr_index_row_bytes')
        return n.value

    @property
    def device_bytes(self):
        """HBM bytes the index holds right now (resident copies + workspaces)."""
        n = ctypes.c_int64(0)
        C.check(C.lib().dhr_index_device_bytes(self._h, ctypes.byref(n)), 'dhr_index_device_bytes')
        return n.value

    def set_option(self, name, value):
        C.check(C.lib().dhr_index_set_option(self._h, name.encode(), int(value)), 'dhr_index_set_option(%s)' % name)

    def stats(self):
        s = C.DhrStats()
        C.check(C.lib().dhr_index_get_stats(self._h, ctypes.byref(s)), 'dhr_index_get_stats')
        return s.as_dict()

    # ---- search ----------------------------------------------------------------------------------
    def _out(self, n, k, out, want_torch):
        if out is not None:
            return out
        if want_torch:
            dev = torch.device('cuda', self.device)
            return (torch.empty((n, k), dtype=torch.float32, device=dev), torch.empty((n, k), dtype=torch.int64, device=dev),
                    torch.empty((n,), dtype=torch.int32, device=dev))
        return (np.empty((n, k), np.float32), np.empty((n, k), np.int64), np.empty((n,), np.int32))

    @staticmethod
    def _ptr(a):
        if torch is not None and isinstance(a, torch.Tensor):
            return a.data_ptr()
        return a.ctypes.data

    def _query_args(self, q_vals, q_idx, masked):
        qv = _Arr(q_vals, 'val')
        if qv.shape[1] != self.width:
            raise ValueError('queries have %d columns, index expects %d' % (qv.shape[1], self.width))
        if self.n_slices > 0 and masked:
            if q_idx is None:
                raise ValueError('index has %d slices: q_idx is required' % self.n_slices)
            qi = _Arr(q_idx, 'idx')
            if qi.shape != (qv.shape[0], self.n_slices):
                raise ValueError('q_idx shape %s, expected %s' % (qi.shape, (qv.shape[0], self.n_slices)))
            return qv, qi, (qi.ptr, qi.code, qi.stride)
        return qv, None, (None, C.IDX_NONE, 0)

    def search(self, q_vals, q_idx, k, lamda=1.0, masked=True, out=None, stream=None, return_torch=None):
        """Top-k rows per query: (scores [Q,k] fp32, rows [Q,k] int64 global ids, counts [Q] int32).

        masked=False is the --IP first stage (gip_retrieval.py:139).  Outputs are numpy arrays unless the
        queries are CUDA tensors (or return_torch=True), in which case they stay on the device."""
        qv, qi, (ip, ic, istr) = self._query_args(q_vals, q_idx, masked)
        n = qv.shape[0]
        want_torch = return_torch if return_torch is not None else (qv.device is not None and qv.device.type == 'cuda')
        scores, rows, counts = self._out(n, k, out, want_torch)
        flags = 0 if masked else C.SEARCH_UNMASKED
        st = ctypes.c_void_p(stream) if stream else None
        C.check(C.lib().dhr_search(self._h, n, qv.code, qv.ptr, qv.stride, ic, ip, istr, float(lamda), int(k), flags,
                                   self._ptr(scores), self._ptr(rows), self._ptr(counts), st), 'dhr_search')
        return scores, rows, counts

    # ---- stream-ordered search (sharded path) ---------------------------------------------------
    def search_keys(self, q_vals, q_idx, k, out_keys, lamda=1.0, masked=True, stream=None):
        """Enqueue the search on `stream` (raw cudaStream_t or None) and return at once: out_keys [Q,k] int64 CUDA tensor
        receives the packed keys (score bits << 32 | 0xFFFFFFFF - global row; descending = (score desc, row asc)).
        Returns (batch_size, n_batches); call wait_batch(b, stream2) / complete() afterwards (include/dhr_b200.h)."""
        qv, qi, (ip, ic, istr) = self._query_args(q_vals, q_idx, masked)
        n = qv.shape[0]
        if tuple(out_keys.shape) != (n, k) or out_keys.dtype != torch.int64 or not out_keys.is_cuda or not out_keys.is_contiguous():
            raise ValueError('out_keys must be a contiguous CUDA int64 tensor of shape [Q, k]')
        flags = 0 if masked else C.SEARCH_UNMASKED
        st = ctypes.c_void_p(stream) if stream else None
        C.check(C.lib().dhr_search_keys(self._h, n, qv.code, qv.ptr, qv.stride, ic, ip, istr, float(lamda), int(k), flags,
                                        out_keys.data_ptr(), st), 'dhr_search_keys')
        bs, nb = ctypes.c_int(0), ctypes.c_int(0)
        C.check(C.lib().dhr_search_batches(self._h, ctypes.byref(bs), ctypes.byref(nb)), 'dhr_search_batches')
        return bs.value, nb.value

    def wait_batch(self, batch, stream):
        C.check(C.lib().dhr_search_wait_batch(self._h, int(batch), ctypes.c_void_p(stream) if stream else None), 'dhr_search_wait_batch')

    def complete(self, stream=None):
        """Synchronise the pending search_keys call; returns the number of queries that were re-run (overflow fallback)."""
        n = ctypes.c_int(0)
        C.check(C.lib().dhr_search_complete(self._h, ctypes.byref(n), ctypes.c_void_p(stream) if stream else None), 'dhr_search_complete')
        return n.value

    def rerank(self, q_vals, q_idx, cand_rows, k, lamda=1.0, out=None, stream=None):
        """Exact GIP over cand_rows [Q, M] (LOCAL row ids, < 0 skipped); same outputs as search()."""
        qv, qi, (ip, ic, istr) = self._query_args(q_vals, q_idx, True)
        n = qv.shape[0]
        if torch is not None and isinstance(cand_rows, torch.Tensor):
            cand = cand_rows.to(torch.int64).contiguous()
        else:
            cand = np.ascontiguousarray(cand_rows, dtype=np.int64)
        if tuple(cand.shape)[0] != n or cand.ndim != 2 if hasattr(cand, 'ndim') else cand.dim() != 2:
            raise ValueError('cand_rows must be [Q, M]')
        want_torch = qv.device is not None and qv.device.type == 'cuda'
        scores, rows, counts = self._out(n, k, out, want_torch)
        st = ctypes.c_void_p(stream) if stream else None
        C.check(C.lib().dhr_rerank(self._h, n, qv.code, qv.ptr, qv.stride, ic, ip, istr, float(lamda), self._ptr(cand),
                                   int(cand.shape[1]), int(k), self._ptr(scores), self._ptr(rows), self._ptr(counts), st),
                'dhr_rerank')
        return scores, rows, counts


def topk_merge(scores, rows, k=None, device=0, stream=None):
    """Merge per-shard top-k lists [P, Q, k] -> ([Q, k] scores, [Q, k] rows) by (score desc, row asc)."""
    is_t = torch is not None and isinstance(scores, torch.Tensor)
    if is_t:
        s = scores.to(torch.float32).contiguous()
        r = rows.to(torch.int64).contiguous()
        P, Q, kk = s.shape
        os_ = torch.empty((Q, kk), dtype=torch.float32, device=s.device)
        or_ = torch.empty((Q, kk), dtype=torch.int64, device=s.device)
        if s.is_cuda:
            device = s.device.index if s.device.index is not None else torch.cuda.current_device()
        ptr = lambda t: t.data_ptr()
    else:
        s = np.ascontiguousarray(scores, dtype=np.float32)
        r = np.ascontiguousarray(rows, dtype=np.int64)
        P, Q, kk = s.shape
        os_ = np.empty((Q, kk), np.float32)
        or_ = np.empty((Q, kk), np.int64)
        ptr = lambda a: a.ctypes.data
    st = ctypes.c_void_p(stream) if stream else None
    C.check(C.lib().dhr_topk_merge(int(device), P, Q, kk, ptr(s), ptr(r), ptr(os_), ptr(or_), st), 'dhr_topk_merge')
    if k is not None and k < kk:
        os_, or_ = os_[:, :k], or_[:, :k]
    return os_, or_


def merge_keys(keys, out_scores=None, out_rows=None, out_keys=None, stream=None):
    """Merge packed-key lists [P, Q, k] (int64 CUDA tensor, each list descending) -> [Q, k]; stream-ordered, no sync.
    Writes (out_scores fp32, out_rows int64) and / or out_keys int64; allocates (scores, rows) when nothing is given."""
    P, Q, k = keys.shape
    if not keys.is_cuda or keys.dtype != torch.int64 or not keys.is_contiguous():
        raise ValueError('keys must be a contiguous CUDA int64 tensor [P, Q, k]')
    if out_scores is None and out_keys is None:
        out_scores = torch.empty((Q, k), dtype=torch.float32, device=keys.device)
        out_rows = torch.empty((Q, k), dtype=torch.int64, device=keys.device)
    ptr = lambda t: None if t is None else t.data_ptr()
    st = ctypes.c_void_p(stream) if stream else None
    C.check(C.lib().dhr_merge_keys(keys.device.index, P, Q, k, keys.data_ptr(), Q * k, ptr(out_scores), ptr(out_rows), ptr(out_keys), st),
            'dhr_merge_keys')
    return out_scores, out_rows, out_keys


def pack_keys(scores, rows):
    """(scores fp32, GLOBAL rows) -> packed int64 keys of the exchange format (host-side mirror of make_key in
    csrc/common.cuh); rows < 0 (padding) -> 0."""
    s = np.ascontiguousarray(scores, dtype=np.float32) + np.float32(0.0)          # -0.0 -> +0.0
    b = s.view(np.uint32)
    o = np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint64)
    r = np.asarray(rows, dtype=np.int64)
    k = (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - r.clip(min=0).astype(np.uint64))
    k[r < 0] = 0
    return k.view(np.int64)


def unpack_keys(keys):
    """packed keys (int64 tensor / array) -> (scores fp32, rows int64); padding (0) -> (-inf, -1).  Host-side helper for tests."""
    a = keys.cpu().numpy() if torch is not None and isinstance(keys, torch.Tensor) else np.asarray(keys)
    u = a.view(np.uint64)
    hi = (u >> np.uint64(32)).astype(np.uint32)
    lo = (u & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    bits = np.where(hi & np.uint32(0x80000000), hi & np.uint32(0x7FFFFFFF), ~hi).astype(np.uint32)
    scores = bits.view(np.float32).copy()
    rows = (np.uint32(0xFFFFFFFF) - lo).astype(np.int64)
    pad = u == 0
    scores[pad] = -np.inf
    rows[pad] = -1
    return scores, rows
