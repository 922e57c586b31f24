"""GPU densify op (next row n3): mirror of castorini/dhr ``tevatron/DHR/utils.py:densify`` that writes the storage
dtypes of the index directly (fp16 values, uint8 slice indices; tevatron/driver/encode.py:155-170)."""
from __future__ import annotations

import ctypes

import torch

from . import _cabi as C


def densify(lexical_reps, dims=768, strategy='stride', remove_dims=570, out=None, stream=None):
    """lexical_reps [batch, vocab] CUDA tensor (fp32 / fp16) -> (value_reps fp16 [batch, dims], index_reps uint8 [batch, dims]).

    Same argument names, checks and error messages as the reference; ``out=(values, idx)`` may be (strided) views into
    the value / index arrays of an index under construction."""
    if not (len(lexical_reps.shape) == 2):
        raise ValueError('Input lexical representation shape should be 2 (batch, vocab), but the input shape is {}'.format(
            len(lexical_reps.shape)))
    orig_dims = lexical_reps.shape[-1]
    if not ((orig_dims - remove_dims) % dims == 0):
        raise ValueError('Input lexical representation cannot be densified, please fix dims or remove_dims')
    if strategy != 'stride':
        raise ValueError('only the stride strategy exists in the reference')
    if not lexical_reps.is_cuda:
        raise ValueError('dhr_b200.densify runs on the GPU: pass a CUDA tensor')
    x = lexical_reps if lexical_reps.stride(1) == 1 else lexical_reps.contiguous()
    code = {torch.float32: C.VAL_F32, torch.float16: C.VAL_F16}.get(x.dtype)
    if code is None:
        x, code = x.float(), C.VAL_F32
    b = x.shape[0]
    if out is None:
        out = (torch.empty((b, dims), dtype=torch.float16, device=x.device), torch.empty((b, dims), dtype=torch.uint8, device=x.device))
    vals, idx = out
    assert vals.dtype == torch.float16 and idx.dtype == torch.uint8 and vals.stride(1) == 1 and idx.stride(1) == 1
    st = ctypes.c_void_p(stream) if stream else None
    C.check(C.lib().dhr_densify(x.device.index or 0, b, orig_dims, dims, remove_dims, code, x.data_ptr(), x.stride(0) if b > 1 else orig_dims,
                                vals.data_ptr(), vals.stride(0) if b > 1 else dims, idx.data_ptr(), idx.stride(0) if b > 1 else dims, st),
            'dhr_densify')
    return vals, idx
