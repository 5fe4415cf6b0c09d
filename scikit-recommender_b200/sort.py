"""`top_k` / `arg_top_k` with the signatures of the reference's native sort module.

Reference: skrec/utils/py/cython/pyx_sort.pyx:104-187 (`pyx_top_k`, `pyx_arg_top_k`; C++ in
include/sort.h:136-170): top-k elements / their indices along the last axis of a 1-D or 2-D array,
sorted by value descending.  Here one CTA per row streams the row once (`k_topk_scores`, the
selection kernel of the score-matrix path) through the C ABI (`skr_topk_scores_host`).

Differences, on purpose: equal values rank the lower index first (the reference's order on ties is a
heap artefact); float32 only, plus int32 whose magnitudes float32 holds exactly
(|x| <= 2^24); `n_threads` is accepted and ignored; k <= 512.  There is no CPU fallback.
"""
import numpy as np

__all__ = ["top_k", "arg_top_k"]

_ctx_cache = {}


def _context(device):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("top_k / arg_top_k need a CUDA device (sm_100a); there is no CPU fallback")
    from . import _native
    dev = torch.cuda.current_device() if device is None else int(device)
    ctx = _ctx_cache.get(dev)
    if ctx is None:
        ctx = _native.Context(dev)
        _ctx_cache[dev] = ctx
    return ctx


def _as_f32_rows(array):
    a = np.asarray(array)
    if a.ndim not in (1, 2):
        raise ValueError("'array' must be 1-dim or 2-dim array_like.")
    if a.dtype == np.int32 or a.dtype == np.int64:
        if a.size and np.max(np.abs(a)) > (1 << 24):
            raise TypeError("integer input beyond 2^24 is not exactly representable in float32")
        src_dtype = a.dtype
    elif a.dtype == np.float32:
        src_dtype = np.float32
    else:
        raise TypeError("The type of 'array' is not supported.")
    rows = np.ascontiguousarray(a.reshape(1, -1) if a.ndim == 1 else a, dtype=np.float32)
    return a.ndim, src_dtype, rows


def arg_top_k(array, top_k, n_threads=1, *, device=None):
    """Indices of the top-k elements along the last axis (int32, value descending)."""
    ndim, _, rows = _as_f32_rows(array)
    idx, _ = _context(device).topk_scores_host(rows, int(top_k), want_idx=True, want_val=False)
    return idx[0] if ndim == 1 else idx


def top_k(array, top_k, n_threads=1, *, device=None):
    """The top-k elements along the last axis (input dtype, descending)."""
    ndim, dt, rows = _as_f32_rows(array)
    _, val = _context(device).topk_scores_host(rows, int(top_k), want_idx=False, want_val=True)
    val = val.astype(dt)
    return val[0] if ndim == 1 else val
