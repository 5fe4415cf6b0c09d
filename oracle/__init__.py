"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see oracle/skr_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  It offers two checkers:

  * the C restatement (`eval_scores`, `topk`, `mask_rows`, `scores`, `mean_f32`, `sums_f64`,
    `evaluate_dicts`), tie policy "score desc, item id asc";
  * the compiled, unmodified reference (`ref_available`, `ref_eval_score_matrix`,
    `ref_evaluate_matrix`, `RefRankingEvaluator`) out of oracle/_ref.
"""
import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libskr_oracle.so")
        if not os.path.exists(path):
            from . import build_ref
            build_ref.build_oracle()
        _LIB = ctypes.CDLL(path)
    return _LIB


def dicts_to_csr(users, d, dedup_sort=False):
    """{user: int array} for `users` (in that order) -> (indptr int64, indices int32)."""
    indptr = np.zeros(len(users) + 1, dtype=np.int64)
    chunks = []
    for r, u in enumerate(users):
        a = np.asarray(d[u], dtype=np.int32).ravel() if (d is not None and u in d) else np.zeros(0, np.int32)
        if dedup_sort:
            a = np.unique(a)
        chunks.append(a)
        indptr[r + 1] = indptr[r] + a.size
    indices = np.concatenate(chunks).astype(np.int32) if chunks else np.zeros(0, np.int32)
    return indptr, np.ascontiguousarray(indices)


def topk(scores_row, k):
    s = np.ascontiguousarray(scores_row, dtype=np.float32)
    idx = np.empty(k, np.int32)
    val = np.empty(k, np.float32)
    rc = lib().skr_oracle_topk_row(_ptr(s, _f32p), ctypes.c_int64(s.size), int(k), _ptr(idx, _i32p), _ptr(val, _f32p))
    if rc:
        raise ValueError("oracle topk failed rc=%d" % rc)
    return idx, val


def mask_rows(scores, indptr, indices):
    assert scores.dtype == np.float32 and scores.flags.c_contiguous
    rc = lib().skr_oracle_mask_rows(_ptr(scores, _f32p), ctypes.c_int64(scores.shape[0]),
                                    ctypes.c_int64(scores.shape[1]), ctypes.c_int64(scores.shape[1]),
                                    _ptr(np.ascontiguousarray(indptr, np.int64), _i64p),
                                    _ptr(np.ascontiguousarray(indices, np.int32), _i32p))
    if rc:
        raise ValueError("oracle mask failed rc=%d" % rc)
    return scores


def eval_scores(scores, test_indptr, test_indices, metric_ids, top_k, return_topk=False):
    """float32 [B, I] (already masked) -> float32 [B, M*K] (+ int32 [B, K])."""
    s = np.ascontiguousarray(scores, dtype=np.float32)
    B, n = s.shape
    m = np.ascontiguousarray(metric_ids, dtype=np.int32)
    out = np.zeros((B, m.size * top_k), np.float32)
    tk = np.empty((B, top_k), np.int32) if return_topk else None
    tp = np.ascontiguousarray(test_indptr, np.int64)
    ti = np.ascontiguousarray(test_indices, np.int32)
    rc = lib().skr_oracle_eval_scores(_ptr(s, _f32p), ctypes.c_int64(B), ctypes.c_int64(n), ctypes.c_int64(n),
                                      _ptr(tp, _i64p), _ptr(ti, _i32p), _ptr(m, _i32p), int(m.size), int(top_k),
                                      _ptr(out, _f32p), _ptr(tk, _i32p))
    if rc:
        raise ValueError("oracle eval_scores failed rc=%d" % rc)
    return (out, tk) if return_topk else out


def scores(user_vecs, item_vecs, bias=None):
    u = np.ascontiguousarray(user_vecs, np.float32)
    v = np.ascontiguousarray(item_vecs, np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    out = np.empty((u.shape[0], v.shape[0]), np.float32)
    lib().skr_oracle_scores(_ptr(u, _f32p), ctypes.c_int64(u.shape[0]), ctypes.c_int64(u.shape[1]),
                            _ptr(v, _f32p), ctypes.c_int64(v.shape[0]), ctypes.c_int64(v.shape[1]), int(u.shape[1]),
                            _ptr(b, _f32p), _ptr(out, _f32p), ctypes.c_int64(v.shape[0]))
    return out


def mean_f32(per_user):
    p = np.ascontiguousarray(per_user, np.float32)
    out = np.empty(p.shape[1], np.float32)
    lib().skr_oracle_mean_f32(_ptr(p, _f32p), ctypes.c_int64(p.shape[0]), ctypes.c_int64(p.shape[1]), _ptr(out, _f32p))
    return out


def sums_f64(per_user):
    p = np.ascontiguousarray(per_user, np.float32)
    out = np.empty(p.shape[1], np.float64)
    lib().skr_oracle_sums_f64(_ptr(p, _f32p), ctypes.c_int64(p.shape[0]), ctypes.c_int64(p.shape[1]), _ptr(out, _f64p))
    return out


def evaluate_dicts(score_fn, user_train, user_test, metric_ids, top_k, users=None, batch=256):
    """evaluator.py:181-208 restated over the C pieces.  `score_fn(users)->float32 [B, I]`.
    Returns (per_user float32 [U, M*K], topk int32 [U, K])."""
    users = list(user_test.keys()) if users is None else [u for u in users if u in user_test]
    outs, tops = [], []
    for b0 in range(0, len(users), batch):
        bu = users[b0:b0 + batch]
        s = np.array(score_fn(bu), dtype=np.float32, copy=True, order="C")
        if user_train:
            tp, ti = dicts_to_csr(bu, user_train)
            mask_rows(s, tp, ti)
        ep, ei = dicts_to_csr(bu, user_test)
        o, t = eval_scores(s, ep, ei, metric_ids, top_k, return_topk=True)
        outs.append(o)
        tops.append(t)
    return np.concatenate(outs, 0), np.concatenate(tops, 0)


# ---------------------------------------------------------------- compiled reference
def ref_dir():
    return os.path.join(_HERE, "_ref")


def ref_available():
    return os.path.exists(os.path.join(ref_dir(), "libref_eval.so"))


def _ref_lib():
    global _REF_LIB
    if _REF_LIB is None:
        _REF_LIB = ctypes.CDLL(os.path.join(ref_dir(), "libref_eval.so"))
    return _REF_LIB


def ref_evaluate_matrix(scores, test_indptr, test_indices, metric_ids, top_k, threads=1):
    """reference cpp_evaluate_matrix (evaluate.h:57-76) through oracle/ref_shim.cpp."""
    s = np.ascontiguousarray(scores, dtype=np.float32)
    B, n = s.shape
    m = np.ascontiguousarray(metric_ids, dtype=np.int32)
    out = np.zeros((B, m.size * top_k), np.float32)
    tp = np.ascontiguousarray(test_indptr, np.int64)
    ti = np.ascontiguousarray(test_indices, np.int32)
    rc = _ref_lib().ref_evaluate_matrix(_ptr(s, _f32p), int(B), int(n), _ptr(tp, _i64p), _ptr(ti, _i32p),
                                        _ptr(m, _i32p), int(m.size), int(top_k), int(threads), _ptr(out, _f32p))
    if rc:
        raise ValueError("reference evaluate_matrix refused rc=%d" % rc)
    return out


def _ref_import():
    d = ref_dir()
    if d not in sys.path:
        sys.path.insert(0, d)


def ref_python_available():
    d = ref_dir()
    return os.path.isdir(os.path.join(d, "refpkg")) and any(
        f.startswith("evaluator.") and f.endswith(".so") for f in os.listdir(os.path.join(d, "refpkg")))


def ref_eval_score_matrix(score_matrix, test_items, metric, top_k, thread_num):
    """the reference's Cython entry point, pyx_eval_matrix.pyx:22-37, unmodified."""
    _ref_import()
    from refpkg.cython import eval_score_matrix
    return eval_score_matrix(score_matrix, test_items, metric, top_k, thread_num)


def ref_module(name):
    """a compiled reference Cython module by name ("pyx_sort", "pyx_eval_matrix", "pyx_utils"), or None"""
    if not ref_python_available():
        return None
    _ref_import()
    import importlib
    try:
        return importlib.import_module("refpkg.cython." + name)
    except ImportError:
        return None


def RefRankingEvaluator(*args, **kwargs):
    """the reference's RankingEvaluator (evaluator.py:61-214), unmodified, compiled by Cython."""
    _ref_import()
    from refpkg.evaluator import RankingEvaluator
    return RankingEvaluator(*args, **kwargs)
