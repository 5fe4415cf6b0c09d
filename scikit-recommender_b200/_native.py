"""ctypes binding of the C ABI in include/skrec_b200.h (libskrec_b200.so, in-tree).

There is no Python/CPU fallback: if the shared library is missing, or no sm_100 device exists,
the product path raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libskrec_b200.so")

SKR_OK = 0
PREC = {"auto": 0, "fp32": 1, "3xtf32": 2, "1xtf32": 3, "tf32r": 4, "f16r": 5}

# every symbol include/skrec_b200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "skr_abi_version", "skr_ctx_create", "skr_ctx_destroy", "skr_last_error", "skr_set_train_csr",
    "skr_set_test_csr", "skr_eval_scores", "skr_eval_scores_host", "skr_eval_fused", "skr_eval_fused_host",
    "skr_metrics_from_topk", "skr_colsum_f32_seq", "skr_launch_count", "skr_last_fused_kernel", "skr_set_option", "skr_fused_kernel_ms", "skr_fused_prepass_ms", "skr_fused_stats", "skr_fused_trace", "skr_topk_fused", "skr_eval_merged_topk", "skr_topk_scores", "skr_topk_scores_host", "skr_colsum_rows",
    "skr_plan_work_host", "skr_comm_create", "skr_comm_handle", "skr_comm_connect", "skr_comm_allreduce", "skr_comm_status",
    "skr_comm_last_error", "skr_comm_destroy", "skr_batch_randint", "skr_check",
)

_lib = None
_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("skrec_b200 native error %d: %s" % (code, message))
        self.code = code


def lib():
    """Load the CUDA extension; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python scikit-recommender_b200/build.py` "
                          "(there is no CPU fallback for the evaluation path)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.skr_abi_version.restype = _int
    L.skr_ctx_create.argtypes = [_int, ctypes.POINTER(_vp)]
    L.skr_ctx_destroy.argtypes = [_vp]
    L.skr_last_error.argtypes = [_vp]
    L.skr_last_error.restype = ctypes.c_char_p
    L.skr_set_train_csr.argtypes = [_vp, _vp, _vp, _i64, _i64]
    L.skr_set_test_csr.argtypes = [_vp, _vp, _vp, _i64, _i64]
    L.skr_eval_scores.argtypes = [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp]
    L.skr_eval_scores_host.argtypes = [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _int, _int, _vp, _vp, _vp, _vp]
    L.skr_eval_fused.argtypes = [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _int, _vp, _i64, _vp, _int, _int, _int,
                                 _vp, _vp, _vp, _vp, _vp]
    L.skr_eval_fused_host.argtypes = [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _int, _vp, _i64, _vp, _int, _int, _int,
                                      _vp, _vp, _vp, _vp]
    L.skr_topk_fused.argtypes = [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _int, _vp, _i64, _i64, _int, _int, _vp, _vp]
    L.skr_eval_merged_topk.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, _i64, _vp, _int, _int, _vp, _vp, _vp, _vp, _vp]
    L.skr_topk_scores.argtypes = [_vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp]
    L.skr_topk_scores_host.argtypes = [_vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp]
    L.skr_colsum_rows.argtypes = [_vp, _vp, _i64, _vp, _i64, _vp, _vp]
    L.skr_metrics_from_topk.argtypes = [_vp, _vp, _i64, _i64, _vp, _int, _int, _vp, _vp, _vp]
    L.skr_colsum_f32_seq.argtypes = [_vp, _vp, _i64, _i64, _vp, _vp]
    L.skr_launch_count.argtypes = [_vp]
    L.skr_launch_count.restype = _i64
    L.skr_last_fused_kernel.argtypes = [_vp]
    L.skr_last_fused_kernel.restype = ctypes.c_char_p
    L.skr_set_option.argtypes = [_vp, ctypes.c_char_p, _i64]
    L.skr_fused_kernel_ms.argtypes = [_vp, _int, ctypes.POINTER(ctypes.c_float)]
    L.skr_fused_prepass_ms.argtypes = [_vp, _int, ctypes.POINTER(ctypes.c_float)]
    L.skr_fused_stats.argtypes = [_vp, ctypes.POINTER(_i64), _int]
    L.skr_fused_trace.argtypes = [_vp, ctypes.POINTER(_i64), _i64]
    L.skr_plan_work_host.argtypes = [_int, _int, _int, _int, _int, _vp, _i64, ctypes.POINTER(_i64)]
    L.skr_plan_work_host.restype = _i64
    L.skr_batch_randint.argtypes = [_vp, _i64, _vp, _i64, _i64, _int, _vp, _int, _vp, _vp, ctypes.c_uint64, _vp, _vp]
    L.skr_check.argtypes = [_vp]
    L.skr_comm_create.argtypes = [_int, _int, _int, ctypes.POINTER(_vp)]
    L.skr_comm_handle.argtypes = [_vp, _vp]
    L.skr_comm_connect.argtypes = [_vp, _vp]
    L.skr_comm_allreduce.argtypes = [_vp, _vp, _int, _vp]
    L.skr_comm_status.argtypes = [_vp]
    L.skr_comm_last_error.argtypes = [_vp]
    L.skr_comm_last_error.restype = ctypes.c_char_p
    L.skr_comm_destroy.argtypes = [_vp]
    for name in SYMBOLS:
        getattr(L, name)
    if L.skr_abi_version() != 1:
        raise ImportError("libskrec_b200.so ABI version %d, expected 1" % L.skr_abi_version())
    _lib = L
    return L


def plan_work(n_user_tiles, n_item_tiles, n_sm=148, cta_overhead=10, chunks=0):
    """Work list of the tcgen05 main pass (host only, no GPU needed): -> (int32 [n, 4] of (user tile, first item tile,
    item tiles, chunk index), dict(slots, min_slots, max_tiles, mixed, makespan))."""
    L = lib()
    info = (_i64 * 5)()
    n = L.skr_plan_work_host(int(n_user_tiles), int(n_item_tiles), int(n_sm), int(cta_overhead), int(chunks), None, 0, info)
    if n < 0:
        raise NativeError(int(n), "skr_plan_work_host: invalid argument")
    items = np.empty((int(n), 4), np.int32)
    L.skr_plan_work_host(int(n_user_tiles), int(n_item_tiles), int(n_sm), int(cta_overhead), int(chunks),
                         ctypes.c_void_p(items.ctypes.data), int(n), info)
    return items, dict(zip(("slots", "min_slots", "max_tiles", "mixed", "makespan"), [int(x) for x in info]))


def _np_ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _dev_ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Context(object):
    """One native context per CUDA device (owns the device CSRs and workspace)."""

    def __init__(self, device=0):
        self._L = lib()
        h = _vp()
        rc = self._L.skr_ctx_create(int(device), ctypes.byref(h))
        if rc != SKR_OK:
            raise NativeError(rc, (self._L.skr_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._L.skr_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != SKR_OK:
            raise NativeError(rc, (self._L.skr_last_error(self._h) or b"").decode())

    # -- CSR ------------------------------------------------------------------------------------
    def set_train_csr(self, indptr, indices, n_items):
        if indptr is None:
            self._check(self._L.skr_set_train_csr(self._h, None, None, 0, int(n_items)))
            return
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        self._check(self._L.skr_set_train_csr(self._h, _np_ptr(ip), _np_ptr(ix), ip.size - 1, int(n_items)))

    def set_test_csr(self, indptr, indices, n_items):
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        self._check(self._L.skr_set_test_csr(self._h, _np_ptr(ip), _np_ptr(ix), ip.size - 1, int(n_items)))

    def set_option(self, name, value):
        self._check(self._L.skr_set_option(self._h, name.encode(), int(value)))

    @property
    def launch_count(self):
        return int(self._L.skr_launch_count(self._h))

    @property
    def last_fused_kernel(self):
        return (self._L.skr_last_fused_kernel(self._h) or b"").decode()

    def fused_kernel_ms(self, back=0):
        ms = ctypes.c_float(0.0)
        self._check(self._L.skr_fused_kernel_ms(self._h, int(back), ctypes.byref(ms)))
        return float(ms.value)

    def fused_prepass_ms(self, back=0):
        ms = ctypes.c_float(0.0)
        self._check(self._L.skr_fused_prepass_ms(self._h, int(back), ctypes.byref(ms)))
        return float(ms.value)

    def fused_stats(self):
        out = (_i64 * 9)()
        self._check(self._L.skr_fused_stats(self._h, out, 9))
        return dict(zip(("sample_tiles", "stride", "rank", "cap", "chunks", "stages", "exact_rows", "timed_launches", "retried_rows"),
                        [int(x) for x in out]))

    def fused_trace(self, n_tiles):
        """[n_tiles, 16] SM-clock timestamps of the traced CTA (set_option("trace_cta", c) first)."""
        import numpy as np
        out = np.zeros((int(n_tiles), 16), dtype=np.int64)
        self._check(self._L.skr_fused_trace(self._h, out.ctypes.data_as(ctypes.POINTER(_i64)), out.size))
        return out

    # -- device entry points (torch CUDA tensors) -----------------------------------------------
    def eval_scores(self, scores, row0, metric_ids, top_k, topk_idx=None, topk_val=None, per_user=None, sums=None,
                    stream=None):
        m = np.ascontiguousarray(metric_ids, dtype=np.int32)
        assert scores.is_cuda and scores.dim() == 2 and scores.stride(1) == 1
        self._check(self._L.skr_eval_scores(self._h, _dev_ptr(scores), scores.shape[0], scores.shape[1], scores.stride(0),
                                            int(row0), _np_ptr(m), int(m.size), int(top_k), _dev_ptr(topk_idx),
                                            _dev_ptr(topk_val), _dev_ptr(per_user), _dev_ptr(sums), _stream_ptr(stream)))

    def eval_fused(self, user_vecs, item_vecs, bias, row0, metric_ids, top_k, precision="auto", topk_idx=None,
                   topk_val=None, per_user=None, sums=None, stream=None):
        m = np.ascontiguousarray(metric_ids, dtype=np.int32)
        assert user_vecs.is_cuda and item_vecs.is_cuda and user_vecs.stride(1) == 1 and item_vecs.stride(1) == 1
        assert user_vecs.shape[1] == item_vecs.shape[1]
        self._check(self._L.skr_eval_fused(self._h, _dev_ptr(user_vecs), user_vecs.shape[0], user_vecs.stride(0),
                                           _dev_ptr(item_vecs), item_vecs.shape[0], item_vecs.stride(0),
                                           int(user_vecs.shape[1]), _dev_ptr(bias), int(row0), _np_ptr(m), int(m.size),
                                           int(top_k), PREC[precision], _dev_ptr(topk_idx), _dev_ptr(topk_val),
                                           _dev_ptr(per_user), _dev_ptr(sums), _stream_ptr(stream)))

    def topk_fused(self, user_vecs, item_vecs, bias, row0, item_offset, top_k, keys_out, precision="auto", stream=None):
        """Sorted top-K rank keys (int64 view of uint64) of the rows over this rank's item shard -> keys_out [n, K]."""
        assert user_vecs.is_cuda and item_vecs.is_cuda and keys_out.is_cuda and keys_out.is_contiguous()
        assert keys_out.shape == (user_vecs.shape[0], int(top_k)) and keys_out.element_size() == 8
        self._check(self._L.skr_topk_fused(self._h, _dev_ptr(user_vecs), user_vecs.shape[0], user_vecs.stride(0),
                                           _dev_ptr(item_vecs), item_vecs.shape[0], item_vecs.stride(0),
                                           int(user_vecs.shape[1]), _dev_ptr(bias), int(row0), int(item_offset), int(top_k),
                                           PREC[precision], _dev_ptr(keys_out), _stream_ptr(stream)))

    def eval_merged_topk(self, keys_all, row_begin, n_rows, row0, metric_ids, top_k, topk_idx=None, topk_val=None,
                         per_user=None, sums=None, stream=None):
        """keys_all: [n_shards, n_rows_total, K] gathered per-shard lists; merge + metrics of a row slice."""
        m = np.ascontiguousarray(metric_ids, dtype=np.int32)
        assert keys_all.is_cuda and keys_all.is_contiguous() and keys_all.dim() == 3 and keys_all.shape[2] == int(top_k)
        self._check(self._L.skr_eval_merged_topk(self._h, _dev_ptr(keys_all), int(keys_all.shape[0]), int(keys_all.shape[1]),
                                                 int(row_begin), int(n_rows), int(row0), _np_ptr(m), int(m.size), int(top_k),
                                                 _dev_ptr(topk_idx), _dev_ptr(topk_val), _dev_ptr(per_user), _dev_ptr(sums),
                                                 _stream_ptr(stream)))

    def topk_scores_host(self, scores, top_k, want_idx=True, want_val=False):
        """Top-k of every row of a host float32 [B, N] block -> (idx int32 [B, k] | None, val float32 [B, k] | None)."""
        s = scores
        assert isinstance(s, np.ndarray) and s.dtype == np.float32 and s.ndim == 2 and s.strides[1] == 4
        idx = np.empty((s.shape[0], int(top_k)), np.int32) if want_idx else None
        val = np.empty((s.shape[0], int(top_k)), np.float32) if want_val else None
        self._check(self._L.skr_topk_scores_host(self._h, _np_ptr(s), s.shape[0], s.shape[1], s.strides[0] // 4, int(top_k),
                                                 _np_ptr(idx), _np_ptr(val), None))
        return idx, val

    def topk_scores(self, scores, top_k, topk_idx=None, topk_val=None, stream=None):
        assert scores.is_cuda and scores.dim() == 2 and scores.stride(1) == 1
        self._check(self._L.skr_topk_scores(self._h, _dev_ptr(scores), scores.shape[0], scores.shape[1], scores.stride(0),
                                            int(top_k), _dev_ptr(topk_idx), _dev_ptr(topk_val), _stream_ptr(stream)))

    def colsum_rows(self, per_user, row_list, sums, stream=None):
        """sums[c] += sum over rows row_list (int32 device tensor) of per_user[:, c] in float64."""
        assert per_user.is_cuda and per_user.is_contiguous() and row_list.is_cuda and sums.is_cuda
        self._check(self._L.skr_colsum_rows(self._h, _dev_ptr(per_user), per_user.shape[1], _dev_ptr(row_list),
                                            int(row_list.numel()), _dev_ptr(sums), _stream_ptr(stream)))

    def metrics_from_topk(self, topk_idx, row0, metric_ids, top_k, per_user=None, sums=None, stream=None):
        m = np.ascontiguousarray(metric_ids, dtype=np.int32)
        self._check(self._L.skr_metrics_from_topk(self._h, _dev_ptr(topk_idx), topk_idx.shape[0], int(row0), _np_ptr(m),
                                                  int(m.size), int(top_k), _dev_ptr(per_user), _dev_ptr(sums),
                                                  _stream_ptr(stream)))

    def colsum_f32_seq(self, per_user, acc, stream=None):
        self._check(self._L.skr_colsum_f32_seq(self._h, _dev_ptr(per_user), per_user.shape[0], per_user.shape[1],
                                               _dev_ptr(acc), _stream_ptr(stream)))

    def batch_randint(self, high, out_indptr, out, replace=True, cdf=None, excl_indptr=None, excl_idx=None, seed=0, stream=None):
        """Device tensors: out_indptr int64 [n + 1], out int32 [out_indptr[-1]], cdf float32 [high] | [n, high] | None,
        exclusion CSR (int64 / int32, rows sorted unique) | None."""
        n = int(out_indptr.numel()) - 1
        per_row = int(cdf is not None and cdf.dim() == 2)
        self._check(self._L.skr_batch_randint(self._h, int(high), _dev_ptr(out_indptr), n, int(out.numel()), int(bool(replace)), _dev_ptr(cdf),
                                              per_row, _dev_ptr(excl_indptr), _dev_ptr(excl_idx), ctypes.c_uint64(int(seed) & (2 ** 64 - 1)),
                                              _dev_ptr(out), _stream_ptr(stream)))

    def check(self):
        self._check(self._L.skr_check(self._h))

    # -- host entry points (numpy arrays; copies inside the call) -------------------------------
    def eval_scores_host(self, scores, row0, metric_ids, top_k, want_topk=False, want_per_user=True):
        s = scores
        assert isinstance(s, np.ndarray) and s.dtype == np.float32 and s.ndim == 2 and s.strides[1] == 4
        m = np.ascontiguousarray(metric_ids, dtype=np.int32)
        B = s.shape[0]
        mk = int(m.size) * int(top_k)
        idx = np.empty((B, top_k), np.int32) if want_topk else None
        pu = np.empty((B, mk), np.float32) if want_per_user else None
        sums = np.zeros(mk, np.float64)
        self._check(self._L.skr_eval_scores_host(self._h, _np_ptr(s), B, s.shape[1], s.strides[0] // 4, int(row0),
                                                 _np_ptr(m), int(m.size), int(top_k), _np_ptr(idx), _np_ptr(pu),
                                                 _np_ptr(sums), None))
        return pu, idx, sums

    def eval_fused_host(self, user_vecs, item_vecs, bias, row0, metric_ids, top_k, precision="auto", want_topk=False,
                        want_per_user=False):
        u, v = user_vecs, item_vecs
        for a in (u, v):
            assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == 2 and a.strides[1] == 4
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        m = np.ascontiguousarray(metric_ids, dtype=np.int32)
        B = u.shape[0]
        mk = int(m.size) * int(top_k)
        idx = np.empty((B, top_k), np.int32) if want_topk else None
        pu = np.empty((B, mk), np.float32) if want_per_user else None
        sums = np.zeros(mk, np.float64)
        self._check(self._L.skr_eval_fused_host(self._h, _np_ptr(u), B, u.strides[0] // 4, _np_ptr(v), v.shape[0],
                                                v.strides[0] // 4, int(u.shape[1]), _np_ptr(b), int(row0), _np_ptr(m),
                                                int(m.size), int(top_k), PREC[precision], _np_ptr(idx), _np_ptr(pu),
                                                _np_ptr(sums), None))
        return pu, idx, sums


class Comm(object):
    """One-shot all-reduce over NVLink peer memory (skr_comm_* of the C ABI): one per (device, process group)."""
    MAX_N = 4096

    def __init__(self, device, rank, world):
        self._L = lib()
        h = _vp()
        rc = self._L.skr_comm_create(int(device), int(rank), int(world), ctypes.byref(h))
        if rc != SKR_OK:
            raise NativeError(rc, (self._L.skr_last_error(None) or b"").decode())
        self._h, self.rank, self.world = h, int(rank), int(world)

    def _check(self, rc):
        if rc != SKR_OK:
            raise NativeError(rc, (self._L.skr_comm_last_error(self._h) or b"").decode())

    def handle(self):
        buf = ctypes.create_string_buffer(64)
        self._check(self._L.skr_comm_handle(self._h, ctypes.cast(buf, _vp)))
        return buf.raw

    def connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        self._check(self._L.skr_comm_connect(self._h, ctypes.cast(ctypes.create_string_buffer(blob, len(blob)), _vp)))

    def allreduce(self, vec, stream=None):
        """in-place SUM over the ranks of a float64 CUDA vector (<= 4096 elements), asynchronous on the stream"""
        assert vec.is_cuda and vec.is_contiguous() and vec.element_size() == 8 and vec.numel() <= self.MAX_N
        self._check(self._L.skr_comm_allreduce(self._h, _dev_ptr(vec), int(vec.numel()), _stream_ptr(stream)))

    def status(self):
        self._check(self._L.skr_comm_status(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._L.skr_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _stream_ptr(stream):
    if stream is None:
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if isinstance(stream, int):
        return ctypes.c_void_p(stream)
    return ctypes.c_void_p(stream.cuda_stream)
