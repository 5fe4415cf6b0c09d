"""`randint_choice` / `batch_randint_choice`: the reference's negative sampler on the GPU.

Reference: skrec/utils/py/random.py:9-41 -> pyx_random.pyx:18-150 -> randint.h:22-128 (callers: data_iterator.py:81-94
`randint_choice(num_items, size=n_pos*num_neg, exclusion=user_pos_dict[user])`, SASRec.py:358, CDAE.py:175).  Same
signatures, argument checks and return types; `thread_num` is accepted and ignored (CUDA grid).  New keyword-only
arguments: `seed` (every draw is a function of (seed, position, attempt) -- Philox4x32-10 -- so a call can be replayed;
default: a fresh seed per call from numpy's global generator), `device`, and `as_tensor=True` to get CUDA tensors back
(`batch_randint_choice`: the flat int32 tensor plus the int64 row pointer) instead of host arrays.

The reference draws from one sequential std::mt19937: its stream is not reproduced (parity is distributional:
range, exclusion, distinctness, uniformity -- tests/test_gpu_parity.py).  There is no CPU fallback.
"""
import numpy as np

__all__ = ["randint_choice", "batch_randint_choice"]

_ctx_cache = {}


def _context(device):
    from . import _native
    ctx = _ctx_cache.get(device)
    if ctx is None:
        ctx = _native.Context(device)
        _ctx_cache[device] = ctx
    return ctx


def _device(device):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("the sampler needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.cuda.current_device() if device is None else int(device)


def _rows_to_csr(rows):
    """list of 1-D int array_like -> (indptr int64, indices int32), rows sorted and unique"""
    rows = [np.unique(np.asarray(list(r) if isinstance(r, (set, frozenset)) else r, dtype=np.int64).ravel()) for r in rows]
    indptr = np.zeros(len(rows) + 1, np.int64)
    np.cumsum([r.size for r in rows], out=indptr[1:])
    idx = np.concatenate(rows).astype(np.int32) if indptr[-1] > 0 else np.zeros(0, np.int32)
    return indptr, idx, rows


def batch_randint_choice(high, size, replace=True, p=None, exclusion=None, thread_num=1, *, seed=None, device=None, as_tensor=False):
    """Sample random integers from [0, high) for every element of a batch (reference random.py:27-41).

    size: 1-D array_like of positive integers; p: 2-D array_like [len(size), high] or None; exclusion: a list of 1-D
    array_like (one per element) or None.  Returns a list of int32 arrays (or, `as_tensor`, (flat CUDA tensor, indptr))."""
    import torch
    if high <= 1:
        raise ValueError("'high' must be larger than 1.")
    if not isinstance(replace, bool):
        raise TypeError("'replace' must be bool.")
    if thread_num < 1 or not isinstance(thread_num, (int, np.int32)):
        raise ValueError("'thread_num' must be a positive integer.")
    try:
        size = np.array(size, np.int32)
    except Exception:
        raise ValueError("'size' must be a 1-dim array_like of positive integers.")
    if size.ndim != 1 or np.any(size <= 0):
        raise ValueError("'size' must be a 1-dim array_like of positive integers.")
    dev_i = _device(device)
    dev = torch.device("cuda", dev_i)
    cdf = None
    if p is not None:
        p = np.asarray(p, dtype=np.float32)
        if p.ndim != 2:
            raise ValueError("'p' must be a 2-dim array_like.")
        if p.shape[0] != len(size):
            raise ValueError("The number of rows of 'p' must be equal with the length of 'size'.")
        if p.shape[1] != high:
            raise ValueError("The number of columns of 'p' must be equal with 'high'.")
        cdf = torch.cumsum(torch.from_numpy(p).to(dev).double(), dim=1).float().contiguous()
    e_ptr = e_idx = None
    if exclusion is not None:
        if len(exclusion) != len(size):
            raise ValueError("The length of 'exclusion' must be equal with the length of 'size'.")
        hp, hi, rows = _rows_to_csr(exclusion)
        for idx, (exc, s) in enumerate(zip(rows, size)):
            if len(exc) >= high:
                raise ValueError("The length of 'exclusion' must be smaller than 'high' in %d-th row." % idx)
            if replace is False and (high - len(exc) <= s):
                raise ValueError("There is not enough integers to be sampled in %d-th row." % idx)
        e_ptr, e_idx = torch.from_numpy(hp).to(dev), torch.from_numpy(hi).to(dev)
    elif replace is False and np.any(high <= size):
        raise ValueError("There is not enough integers to be sampled.")
    indptr = np.zeros(len(size) + 1, np.int64)
    np.cumsum(size, out=indptr[1:])
    d_ptr = torch.from_numpy(indptr).to(dev)
    out = torch.empty(int(indptr[-1]), dtype=torch.int32, device=dev)
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 62))
    ctx = _context(dev_i)
    with torch.cuda.device(dev):
        ctx.batch_randint(high, d_ptr, out, replace=replace, cdf=cdf, excl_indptr=e_ptr, excl_idx=e_idx, seed=seed)
        if as_tensor:
            return out, d_ptr
        host = out.cpu().numpy()
        ctx.check()
    return [host[indptr[i]:indptr[i + 1]] for i in range(len(size))]


def randint_choice(high, size=1, replace=True, p=None, exclusion=None, *, seed=None, device=None):
    """Sample random integers from [0, high) (reference random.py:9-24): int when size == 1, else an int32 ndarray."""
    if high <= 1:
        raise ValueError("'high' must be larger than 1.")
    if size <= 0:
        raise ValueError("'size' must be a positive integer.")
    if not isinstance(replace, bool):
        raise TypeError("'replace' must be bool.")
    if p is not None:
        p = np.asarray(p, dtype=np.float32)
        if p.ndim != 1:
            raise ValueError("'p' must be a 1-dim array_like")
        if len(p) != high:
            raise ValueError("The length of 'p' must be equal with 'high'.")
        p = p[None, :]
    if isinstance(exclusion, (int, np.integer)):
        exclusion = [exclusion]
    if exclusion is not None and len(exclusion) >= high:
        raise ValueError("The length of 'exclusion' must be smaller than 'high'.")
    len_exclusion = len(exclusion) if exclusion is not None else 0
    if replace is False and (high - len_exclusion <= size):
        raise ValueError("There is not enough integers to be sampled.")
    res = batch_randint_choice(high, [int(size)], replace=replace, p=p, exclusion=None if exclusion is None else [exclusion],
                               seed=seed, device=device)[0]
    return res[0] if len(res) == 1 else res
