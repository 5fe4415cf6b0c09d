"""`eval_score_matrix`: drop-in for the reference's native entry point.

Reference: skrec/utils/py/cython/pyx_eval_matrix.pyx:22-37 (callers: evaluator.py:202-203,
BERT4Rec's bert4rec_utils.py:78-79).  Same signature and return value -- a float32
[len(test_items), len(metric) * top_k] array laid out [metric0@1..K | metric1@1..K | ...] -- but
the work is done by the CUDA score-matrix kernels through the C ABI (skr_eval_scores_host).
Differences, on purpose: the input is validated/converted to float32 C-order instead of being
silently misread; ties rank the lower item id first.
"""
import numpy as np

from .evaluator import _as_i32

_ctx_cache = {}


def _context(device):
    from . import _native
    ctx = _ctx_cache.get(device)
    if ctx is None:
        ctx = _native.Context(device)
        _ctx_cache[device] = ctx
    return ctx


def eval_score_matrix(score_matrix, test_items, metric, top_k, thread_num=None, *, device=None):
    """score_matrix: float32 [B, N] with train items already set to -inf by the caller;
    test_items: list of B int arrays; metric: list of ids 1..5; thread_num: ignored (CUDA grid)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("eval_score_matrix needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    s = np.ascontiguousarray(score_matrix, dtype=np.float32)
    if s.ndim != 2:
        raise ValueError("score_matrix must be 2-D")
    if len(test_items) != s.shape[0]:
        raise ValueError("len(test_items) != number of score rows")
    indptr = np.zeros(len(test_items) + 1, dtype=np.int64)
    np.cumsum([len(t) for t in test_items], out=indptr[1:])
    # rows may be arrays, lists or Python sets (BERT4Rec passes `set(items[-1:])`, bert4rec_utils.py:25; the
    # reference converts them through Cython's cset[int], pyx_eval_matrix.pyx:27)
    indices = (np.concatenate([_as_i32(t) for t in test_items]) if indptr[-1] > 0 else np.zeros(0, np.int32))
    ctx = _context(dev)
    ctx.set_train_csr(None, None, s.shape[1])
    ctx.set_test_csr(indptr, indices, s.shape[1])
    per_user, _, _ = ctx.eval_scores_host(s, 0, list(metric), int(top_k), want_topk=False, want_per_user=True)
    return per_user
