"""Host-side mirror of scikit-recommender's evaluation API over the B200 kernels.

Drop-in for `skrec/utils/py/evaluator.py` of the reference: same class names, constructor
signature and defaults (evaluator.py:82-86), public attributes (`metrics`, `metrics_num`,
`max_top`, `top_show`, `num_thread`, `batch_size`, `user_pos_train`, `user_pos_test`; read
directly by bert4rec_utils.py:31-48), exceptions (evaluator.py:118,121,144,180,193) and result
strings (evaluator.py:25-43,151-161).  `evaluate()` (evaluator.py:163-214) keeps its contract --
users = keys of the test dict (or the filtered `test_users`), train items masked, top-K,
Precision/Recall/MAP/NDCG/MRR@1..max_top averaged over users, columns `top_show` reported -- but
its body runs on the GPU:

  * fused path: a model that offers `eval_embeddings(users) -> (user_vecs[B,d], item_vecs[I,d],
    bias[I] | None)` (the operands of its own `predict`: BPRMF.py:84-88, LightGCN.py:102-107,
    MultVAE.py:138-141, SelfCF.py:235-241) is scored tile by tile on tensor cores; the B x I score
    matrix is never materialised.
  * score-matrix path: any other model is asked for `predict(batch_users)` exactly like the
    reference (evaluator.py:192); the returned block is uploaded and masked/ranked on device.

New arguments are keyword-only with defaults, so reference call sites (base.py:25-29) work
unchanged.  There is no CPU fallback.
"""
__all__ = ["MetricReport", "RankingEvaluator", "EarlyStopping"]

from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from .report import EarlyStopping, MetricReport, colour_join

class _nvtx(object):
    """NVTX range around a phase of an evaluate (visible in Nsight Systems / ncu --nvtx); nothing without CUDA."""

    def __init__(self, name):
        self.name, self.on = name, False

    def __enter__(self):
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.nvtx.range_push(self.name)
                self.on = True
        except Exception:  # pragma: no cover - NVTX is best effort
            self.on = False
        return self

    def __exit__(self, *exc):
        if self.on:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


_metric2id = {"Precision": 1, "Recall": 2, "MAP": 3, "NDCG": 4, "MRR": 5}
_id2metric = {value: key for key, value in _metric2id.items()}


def _as_i32(items):
    """One user's items -> flat int32 array.  The reference accepts arrays, lists and (through Cython's `cset[int]`
    conversion, pyx_eval_matrix.pyx:27; BERT4Rec passes `set(items[-1:])`, bert4rec_utils.py:25) Python sets."""
    if isinstance(items, (set, frozenset)):
        return np.fromiter(items, dtype=np.int32, count=len(items))
    return np.asarray(items, dtype=np.int32).ravel()


def _dict_to_csr(users, d, col_range=None):
    """{user: int array} restricted to `users` (in order) -> (indptr int64, indices int32).
    col_range=(lo, hi): keep only items in [lo, hi) and renumber them from 0 (an item shard's column
    partition of the interactions)."""
    counts = np.fromiter((len(d[u]) if u in d else 0 for u in users), dtype=np.int64, count=len(users))
    indptr = np.zeros(len(users) + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    if indptr[-1] == 0:
        return indptr, np.zeros(0, np.int32)
    indices = np.concatenate([_as_i32(d[u]) for u in users if u in d and len(d[u]) > 0])
    if col_range is not None:
        lo, hi = col_range
        keep = (indices >= lo) & (indices < hi)
        rows = np.repeat(np.arange(len(users), dtype=np.int64), counts)
        kept = np.bincount(rows[keep], minlength=len(users)).astype(np.int64)
        indptr = np.zeros(len(users) + 1, dtype=np.int64)
        np.cumsum(kept, out=indptr[1:])
        indices = indices[keep] - np.int32(lo)
    return indptr, np.ascontiguousarray(indices, dtype=np.int32)


class _LazyRows(object):
    """Read-only {user: items} view of a CSR (what code that pokes at `user_pos_train` / `user_pos_test`
    expects) without materialising a dict of arrays."""

    def __init__(self, indptr, indices, keys=None):
        self.indptr, self.indices = indptr, indices
        self._keys = keys  # None: every row

    def __len__(self):
        return int(self.indptr.size - 1) if self._keys is None else int(self._keys.size)

    def __contains__(self, u):
        u = int(u)
        return 0 <= u < self.indptr.size - 1 and (self._keys is None or self.indptr[u + 1] > self.indptr[u])

    def __getitem__(self, u):
        if u not in self:
            raise KeyError(u)
        return self.indices[self.indptr[u]:self.indptr[u + 1]]

    def keys(self):
        return (range(self.indptr.size - 1) if self._keys is None else self._keys.tolist())

    def __iter__(self):
        return iter(self.keys())

    def items(self):
        return ((u, self[u]) for u in self.keys())


def _csr_rows(indptr, indices, rows, col_range=None):
    """rows of a CSR, in the given order -> (indptr int64, indices int32), vectorised"""
    rows = np.asarray(rows, dtype=np.int64)
    if col_range is None and rows.size > 0 and int(rows[-1]) - int(rows[0]) + 1 == rows.size and \
            (rows.size < 3 or bool(np.all(np.diff(rows) == 1))):
        # a run of consecutive rows (every user of a from_csr evaluator, or a rank's contiguous slice): views, no gather
        lo, hi = int(rows[0]), int(rows[-1]) + 1
        return indptr[lo:hi + 1] - indptr[lo], indices[indptr[lo]:indptr[hi]]
    cnt = indptr[rows + 1] - indptr[rows]
    optr = np.zeros(rows.size + 1, dtype=np.int64)
    np.cumsum(cnt, out=optr[1:])
    if optr[-1] == 0:
        return optr, np.zeros(0, np.int32)
    src = np.repeat(indptr[rows] - optr[:-1], cnt) + np.arange(optr[-1], dtype=np.int64)
    oidx = indices[src]
    if col_range is not None:
        lo, hi = col_range
        keep = (oidx >= lo) & (oidx < hi)
        kept = np.bincount(np.repeat(np.arange(rows.size, dtype=np.int64), cnt)[keep], minlength=rows.size).astype(np.int64)
        optr = np.zeros(rows.size + 1, dtype=np.int64)
        np.cumsum(kept, out=optr[1:])
        oidx = oidx[keep] - np.int32(lo)
    return optr, np.ascontiguousarray(oidx, dtype=np.int32)


def _pairs_to_csr(pairs, num_users, num_items):
    """(user, item) pairs [n, 2] -> (indptr int64 [num_users + 1], indices int32), items in the pairs' order per user."""
    users = np.ascontiguousarray(pairs[:, 0], dtype=np.int64)
    items = np.ascontiguousarray(pairs[:, 1], dtype=np.int64)
    if users.size and (users.max() >= num_users or items.max() >= num_items):
        raise ValueError("interaction pair outside [0, %d) x [0, %d)" % (num_users, num_items))
    indptr = np.zeros(int(num_users) + 1, dtype=np.int64)
    np.cumsum(np.bincount(users, minlength=int(num_users)), out=indptr[1:])
    order = np.argsort(users, kind="stable")
    return indptr, items[order].astype(np.int32)


def _read_pairs(path, sep="\t"):
    """First two columns of a header-less interaction file -> int64 [n, 2] (the files `_read_csv` loads, dataset.py:36-42)."""
    import os
    if not os.path.isfile(path):
        raise FileNotFoundError("'%s' does not exist." % path)  # dataset.py:389-395 (`handle=raise_error`)
    import pandas as pd
    df = pd.read_csv(path, sep=sep, header=None, usecols=[0, 1])
    if df.isnull().values.any():
        raise ValueError("'%s' has empty fields, please check the file or the separator." % path)
    return df.to_numpy(dtype=np.int64)


def _rows_to_csr(users, d, col_range=None):
    if isinstance(d, _LazyRows):
        return _csr_rows(d.indptr, d.indices, users, col_range)
    return _dict_to_csr(users, d, col_range)


class _Plan(object):
    """Device state for one evaluated-user list: native context + CSRs of exactly those rows."""

    def __init__(self, users, n_items, train, test, device, item_range=None):
        """item_range=(lo, hi): this rank scores only item rows [lo, hi) (item-sharded evaluation): the
        train CSR becomes that column partition with shard-local ids; the test CSR keeps global ids."""
        import time
        from . import _native
        t0 = time.perf_counter()
        self.users = users
        self.users_arr = np.asarray(users, dtype=np.int64)
        self.n_items = n_items
        self.item_range = item_range
        self.ctx = _native.Context(device)
        tp, ti = _rows_to_csr(users, test)
        self.ctx.set_test_csr(tp, ti, n_items)
        n_local = n_items if item_range is None else item_range[1] - item_range[0]
        if train is not None and len(train) > 0:
            rp, ri = _rows_to_csr(users, train, item_range)
            self.ctx.set_train_csr(rp, ri, n_local)
        else:
            self.ctx.set_train_csr(None, None, n_local)
        self.build_ms = (time.perf_counter() - t0) * 1e3  # CSR slicing + normalisation + mask keys + upload

    def close(self):
        """Free the native context (device CSRs and workspace) now rather than when the object is collected."""
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None


class RankingEvaluator(object):
    """Evaluator for item ranking task (reference evaluator.py:61-214), GPU-resident.

    Args (reference): user_train_dict, user_test_dict, metric, top_k, batch_size, num_thread.
        `num_thread` is accepted and stored (bert4rec_utils.py:79 reads it) but unused: the CUDA
        grid replaces the thread pool.  `batch_size` is the user batch of the `predict` path.
    Keyword-only additions:
        device: CUDA device index (default: current torch device).
        precision: "auto" | "3xtf32" | "fp32" | "tf32r" | "f16r" -- arithmetic of the fused scoring
            ("tf32r" / "f16r": one TF32 / scaled-FP16 pass finds candidates inside a rigorous error band, the survivors are
            re-scored in exact FP32; same results as "fp32").
        mean: "f64" (float64 sums, rounded once to float32) or "numpy_f32" (the reference's
            float32 row-order accumulation of np.mean, evaluator.py:208, bit for bit;
            single-process only).
        shard_users: opt-in.  With torch.distributed initialised, each rank evaluates a contiguous slice
            of the users and the metric sums are all-reduced: `evaluate` becomes a COLLECTIVE that every
            rank of `process_group` must call.  The default (False) keeps reference-style call sites such
            as `if rank == 0: evaluator.evaluate(model)` working unchanged under DDP.
        shard: "users" (default; the item table is replicated) or "items" (catalogue beyond one HBM:
            every rank holds a contiguous range of item rows, computes every user's top-K over its
            range, the per-rank lists are all-gathered and merged, then the sums all-reduced).  With
            "items" the model's `eval_embeddings` may take `item_shard=(rank, world)` and return just
            its rows plus the catalogue size, `(user_vecs, item_rows, bias_rows | None, n_items)`.
        process_group: the group to reduce over (default: WORLD).
        upload: how user-sharded ranks bring a HOST item table to their GPUs.  "sharded" (default): the table is
            taken to be identical on every rank (it is the replicated model); each rank uploads 1/world of its rows
            over PCIe and the slices are all-gathered over NVLink.  "replicated": every rank uploads the whole table.
    """

    def __init__(self, user_train_dict: Optional[Dict[int, np.ndarray]],
                 user_test_dict: Dict[int, np.ndarray],
                 metric: Union[None, str, Tuple[str], List[str]] = None,
                 top_k: Union[int, List[int], Tuple[int]] = 50,
                 batch_size: int = 256, num_thread: int = 8, *,
                 device: Optional[int] = None, precision: str = "auto", mean: str = "f64",
                 shard_users: bool = False, shard: str = "users", process_group=None, upload: str = "sharded"):
        super(RankingEvaluator, self).__init__()
        if metric is None:
            metric = ["Precision", "Recall", "MAP", "NDCG", "MRR"]
        elif isinstance(metric, str):
            metric = [metric]
        elif isinstance(metric, (tuple, list)):
            metric = list(metric)
        else:
            raise TypeError("The type of 'metric' (%s) is invalid!" % metric.__class__.__name__)

        for m in metric:
            assert m in _metric2id, f"'{metric}' is not in ('Precision', 'Recall', 'MAP', 'NDCG', 'MRR')."

        self.user_pos_train = dict()
        self.user_pos_test = dict()
        self._plans = OrderedDict()
        self._slices = {}
        self.set_train_data(user_train_dict)
        self.set_test_data(user_test_dict)

        self.metrics_num = len(metric)
        self.metrics = [_metric2id[m] for m in metric]
        self.num_thread = num_thread
        self.batch_size = batch_size

        if isinstance(top_k, int):
            self.max_top = top_k
            self.top_show = np.arange(top_k) + 1
        else:
            self.max_top = max(top_k)
            self.top_show = np.sort(top_k)

        import os
        allowed = ("auto", "3xtf32", "fp32", "tf32r", "f16r") + (("1xtf32",) if os.environ.get("SKR_ALLOW_1XTF32") == "1" else ())
        # "1xtf32" (one TF32 pass, no re-scoring) misses the 1e-5 metric contract: measurement only,
        # reachable only with SKR_ALLOW_1XTF32=1
        assert precision in allowed, "precision must be auto|3xtf32|fp32|tf32r|f16r"
        assert mean in ("f64", "numpy_f32"), "mean must be f64|numpy_f32"
        assert shard in ("users", "items"), "shard must be users|items"
        assert upload in ("sharded", "replicated"), "upload must be sharded|replicated"
        self.upload = upload
        self.shard = shard
        self.device = device
        self.precision = precision
        self.mean = mean
        self.shard_users = shard_users
        self.process_group = process_group
        self.item_chunk_rows = 1 << 18  # shard='items': users per top-K / all-gather / merge round
        self.last_stats = {}

    @classmethod
    def from_csr(cls, train_csr, test_csr, **kwargs):
        """Build the evaluator straight from interaction matrices, bypassing dict-of-arrays.

        `train_csr` / `test_csr`: scipy.sparse CSR matrices [num_users, num_items] (what
        `ImplicitFeedback.to_csr_matrix()` gives, dataset.py:131-156) or `(indptr, indices)` pairs;
        `train_csr` may be None.  Evaluated users = rows with at least one test item, ascending -- the
        order `to_user_dict()` produces (dataset.py:153-155).  At 10^6 users the reference's dicts of
        small arrays dominate set-up time; here nothing is built per user in Python."""
        def split(m):
            if m is None:
                return None
            if hasattr(m, "indptr") and hasattr(m, "indices"):
                return np.asarray(m.indptr, dtype=np.int64), np.asarray(m.indices, dtype=np.int32)
            ip, ix = m
            return np.asarray(ip, dtype=np.int64), np.asarray(ix, dtype=np.int32)
        tr, te = split(train_csr), split(test_csr)
        users = np.flatnonzero(np.diff(te[0]) > 0)
        assert users.size > 0, "'test_csr' can be empty."
        self = cls(None, _LazyRows(te[0], te[1], users), **kwargs)
        self.user_pos_train = _LazyRows(tr[0], tr[1], None) if tr is not None else dict()
        self._csr = (tr, te)
        return self

    @classmethod
    def from_pairs(cls, train_pairs, test_pairs, num_users=None, num_items=None, **kwargs):
        """Build the evaluator from (user, item) interaction pairs -- int arrays [n, 2], what
        `ImplicitFeedback.to_user_item_pairs()` gives (dataset.py:117-120); `train_pairs` may be None.
        Rows keep the pairs' order per user, like `to_user_dict()` (dataset.py:148-156: groupby user, items in
        file order); users and items are the ids as they come (already remapped by the reference's preprocessor).
        `num_users` / `num_items` default to max id + 1 over both sets (dataset.py:407-411)."""
        tr = None if train_pairs is None else np.asarray(train_pairs).reshape(-1, 2)
        te = np.asarray(test_pairs).reshape(-1, 2)
        assert te.shape[0] > 0, "'user_test_dict' can be empty."
        both = [x for x in (tr, te) if x is not None and x.shape[0] > 0]
        lo = min(int(x.min()) for x in both)
        if lo < 0:
            raise ValueError("negative user or item id in the interaction pairs")
        if num_users is None:
            num_users = max(int(x[:, 0].max()) for x in both) + 1
        if num_items is None:
            num_items = max(int(x[:, 1].max()) for x in both) + 1
        return cls.from_csr(None if tr is None else _pairs_to_csr(tr, num_users, num_items),
                            _pairs_to_csr(te, num_users, num_items), **kwargs)

    @classmethod
    def from_files(cls, train_file, test_file, sep="\t", num_users=None, num_items=None, **kwargs):
        """Build the evaluator straight from the reference's `<prefix>.train` / `<prefix>.test` interaction files
        (dataset.py:388-395: header-less, `sep`-separated, columns user, item[, rating][, time] with integer ids),
        bypassing the DataFrame -> ImplicitFeedback -> dict-of-arrays -> pickle-cache chain (dataset.py:131-156,
        300-362).  Only the first two columns are read.  `train_file` may be None."""
        tr = None if train_file is None else _read_pairs(train_file, sep)
        return cls.from_pairs(tr, _read_pairs(test_file, sep), num_users=num_users, num_items=num_items, **kwargs)

    def set_train_data(self, user_train_dict: Optional[Dict[int, np.ndarray]] = None):
        """Replace the train interactions.  The device copies are rebuilt on the next evaluate; editing the dict's
        arrays in place afterwards is not seen (the reference re-reads the dicts every call) -- call this again."""
        self.user_pos_train = user_train_dict if user_train_dict is not None else dict()
        self._drop_plans()

    def set_test_data(self, user_test_dict: Dict[int, np.ndarray]):
        assert len(user_test_dict) > 0, "'user_test_dict' can be empty."
        self.user_pos_test = user_test_dict
        self._all_users = list(user_test_dict.keys())  # evaluation order of evaluate(model) (evaluator.py:184)
        self._drop_plans()

    @property
    def metrics_list(self) -> List[str]:
        return [f"{_id2metric[mid]}@{str(k)}" for mid in self.metrics for k in self.top_show]

    @property
    def metrics_str(self) -> str:
        """All metric names, coloured and tab-joined (reference evaluator.py:151-161)."""
        return colour_join(self.metrics_list)

    # ------------------------------------------------------------------------------------------
    def _device_index(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("RankingEvaluator needs a CUDA device (sm_100a); there is no CPU fallback")
        return torch.cuda.current_device() if self.device is None else int(self.device)

    _MAX_PLANS = 3  # every plan owns a native context (device CSRs + grow-only workspace: GBs at 10^6 users)

    def _drop_plans(self):
        for plan in getattr(self, "_plans", {}).values():
            plan.close()
        self._plans = OrderedDict()
        self._slices = {}

    def _plan(self, users, n_items, key, item_range=None):
        """Device state for this user list.  A cached plan is reused only if it was built for exactly these users
        (the key of a `test_users` subset is a hash: equality is checked, not assumed)."""
        plan = self._plans.get(key)
        if plan is not None and plan.n_items == n_items and plan.item_range == item_range and \
                (plan.users is users or np.array_equal(plan.users_arr, np.asarray(users, dtype=np.int64))):
            self._plans.move_to_end(key)
            return plan
        if plan is not None:
            del self._plans[key]
            plan.close()
        with _nvtx("skrec:plan (CSR upload)"):
            plan = _Plan(users, n_items, self.user_pos_train, self.user_pos_test, self._device_index(), item_range)
        self._plans[key] = plan
        while len(self._plans) > self._MAX_PLANS:
            _, old = self._plans.popitem(last=False)
            old.close()
        return plan

    def _shard(self, n):
        """(rank, world, lo, hi): this rank's contiguous slice of n evaluated users."""
        from . import dist
        if not self.shard_users and self.shard != "items":
            return 0, 1, 0, n
        rank, world = dist.rank_world(self.process_group)
        lo, hi = dist.shard_range(n, rank, world)
        return rank, world, lo, hi

    def _resolve_users(self, test_users):
        """-> (evaluated users in order, cache key).  evaluator.py:181-184: all test users in dict order, or
        `test_users` filtered to those with test items, in the caller's order."""
        if test_users is not None:
            assert isinstance(test_users, Iterable), "'test_user' must be iterable."
            test_users = [u for u in test_users if u in self.user_pos_test]
            arr = np.asarray(test_users, dtype=np.int64)
            return test_users, ("subset", len(test_users), hash(arr.tobytes()))
        if len(self._all_users) != len(self.user_pos_test):  # the dict was mutated behind our back
            self._all_users = list(self.user_pos_test.keys())
            self._drop_plans()
        return self._all_users, ("all",)

    def _evaluate_packed(self, model, test_users, host_fast):
        """The evaluation up to the exchanged vector.  -> dict(packed, host_sums, n_users, per_user, key, path, world, MK)

        packed: float64 device tensor [M*K + 1] = [column sums | user count], summed over the ranks when sharded --
        everything stays on the device and on the current stream (kernels, then one NCCL all-reduce); nothing is
        synchronised.  host_fast: single-process evaluation of HOST tables may instead run as one native call that
        does its own copies and returns `host_sums` (float64 numpy [M*K]); `packed` is then None."""
        import torch
        from . import dist

        assert hasattr(model, "predict") or hasattr(model, "eval_embeddings"), "the model must have attribute 'predict'."
        users_all, key_all = self._resolve_users(test_users)
        rank, world, lo, hi = self._shard(len(users_all))
        item_sharded = self.shard == "items" and world > 1
        # argument combinations are checked before any work or communication
        if self.mean == "numpy_f32" and world > 1:
            raise RuntimeError("mean='numpy_f32' reproduces a sequential sum and is single-process only")
        if item_sharded:
            assert hasattr(model, "eval_embeddings"), "shard='items' needs a model with eval_embeddings"
            assert self.mean == "f64", "shard='items' supports mean='f64' only"
            lo, hi = 0, len(users_all)  # every rank sees every user; the items are what is split
        key = key_all + (rank, world, self.shard)
        if world > 1 and not item_sharded:
            users = self._slices.get(key) if key_all == ("all",) else None
            if users is None:
                users = users_all[lo:hi]
                if key_all == ("all",):
                    self._slices[key] = users
        else:
            users = users_all
        dev = torch.device("cuda", self._device_index())
        K, M = self.max_top, self.metrics_num
        MK = M * K
        want_pu = self.mean == "numpy_f32"
        out = dict(packed=None, host_sums=None, n_users=len(users), per_user=None, key=key, path="none", world=world,
                   MK=MK, dev=dev)
        with torch.cuda.device(dev):
            packed = torch.zeros(MK + 1, dtype=torch.float64, device=dev)
            sums = packed[:MK]
            # a rank left without users (fewer users than ranks) still takes part in the item table's all-gather
            gather_only = len(users) == 0 and world > 1 and not item_sharded and self.upload == "sharded" and \
                hasattr(model, "eval_embeddings")
            if len(users) > 0 or gather_only:
                if item_sharded:
                    out["path"], out["n_users"] = self._evaluate_item_sharded(model, users, key, dev, rank, world, sums)
                elif hasattr(model, "eval_embeddings"):
                    out["path"], out["host_sums"], out["per_user"] = self._evaluate_fused(
                        model, users, key, dev, want_pu, sums, host_fast and world == 1, rank, world)
                else:
                    out["path"], out["per_user"] = self._evaluate_predict(model, users, key, dev, want_pu, sums)
            if out["host_sums"] is None:
                packed[MK:].fill_(float(out["n_users"]))
                if world > 1:
                    with _nvtx("skrec:allreduce [sums | count]"):
                        dist.allreduce_sums(packed, self.process_group)
                out["packed"] = packed
        self.last_stats = {"path": out["path"], "users": out["n_users"], "world": world}
        return out

    def last_context(self):
        """The native context of the most recently used plan (instrumentation: launch counts, kernel timings), or None."""
        if not self._plans:
            return None
        return next(reversed(self._plans.values())).ctx

    def evaluate_device(self, model, test_users: Optional[Iterable[int]] = None):
        """`evaluate` without its final device-to-host copy: -> float64 CUDA tensor [n_metrics * max_top + 1] holding
        [column sums over all evaluated users | number of users] (already all-reduced when sharded), produced
        asynchronously on the current stream.  `evaluate` is this plus one copy to the host and the division."""
        return self._evaluate_packed(model, test_users, host_fast=False)["packed"]

    def evaluate(self, model, test_users: Optional[Iterable[int]] = None) -> MetricReport:
        """Evaluate `model` (reference evaluator.py:163-214).

        `model` must have `predict(users) -> float32 ndarray [B, num_items]`; if it also has
        `eval_embeddings(users)` the fused path is used.
        """
        import torch
        from . import dist

        r = self._evaluate_packed(model, test_users, host_fast=True)
        MK = r["MK"]
        if self.mean == "numpy_f32":
            if r["per_user"] is None:  # nobody to evaluate
                final_results = np.zeros(MK, np.float32)
            else:
                plan = self._plans[r["key"]]
                with torch.cuda.device(r["dev"]):
                    acc = torch.zeros(MK, dtype=torch.float32, device=r["dev"])
                    plan.ctx.colsum_f32_seq(r["per_user"], acc)
                    final_results = (acc / torch.tensor(float(r["n_users"]), dtype=torch.float32, device=r["dev"])).cpu().numpy()
        elif r["host_sums"] is not None:
            final_results = dist.finalize_means(r["host_sums"], r["n_users"])
        else:
            host = r["packed"].cpu().numpy()  # the one device-to-host copy (and synchronisation) of an evaluate
            final_results = dist.finalize_means(host[:MK], host[MK])

        final_results = np.reshape(final_results, [self.metrics_num, self.max_top])
        final_results = final_results[:, self.top_show - 1]
        final_results = np.reshape(final_results, [-1])
        return MetricReport(self.metrics_list, final_results)

    def evaluate_groups(self, model, groups) -> List[MetricReport]:
        """One evaluation for several user groups (reference: `evaluate_group`, base.py:66-71, runs
        `evaluate(group.users)` once per group -- 4 extra full passes for the activity groups of
        dataset.py:707-765).  `groups`: iterables of user ids.  The union of the groups' test users is
        evaluated once with the per-user metric block kept on the device; each group's report is the
        float64 mean of its users' rows.  Same numbers as `[evaluate(model, g) for g in groups]`."""
        import torch
        groups = [[u for u in g if u in self.user_pos_test] for g in groups]
        union, pos = [], {}
        for g in groups:
            for u in g:
                if u not in pos:
                    pos[u] = len(union)
                    union.append(u)
        K, M = self.max_top, self.metrics_num
        MK = M * K
        if not union:
            return [MetricReport(self.metrics_list, np.zeros(M * len(self.top_show), np.float32)) for _ in groups]
        dev = torch.device("cuda", self._device_index())
        key = ("subset", len(union), hash(np.asarray(union, dtype=np.int64).tobytes()), 0, 1, "groups")
        with torch.cuda.device(dev):
            scratch = torch.zeros(MK, dtype=torch.float64, device=dev)
            if hasattr(model, "eval_embeddings"):
                _, _, per_user = self._evaluate_fused(model, union, key, dev, True, scratch, False)
            else:
                _, per_user = self._evaluate_predict(model, union, key, dev, True, scratch)
            plan = self._plans[key]
            sums = torch.zeros((len(groups), MK), dtype=torch.float64, device=dev)
            for gi, g in enumerate(groups):
                if g:
                    rows = torch.from_numpy(np.fromiter((pos[u] for u in g), dtype=np.int32, count=len(g))).to(dev)
                    plan.ctx.colsum_rows(per_user, rows, sums[gi])
            host = sums.cpu().numpy()
        from . import dist
        out = []
        for gi, g in enumerate(groups):
            res = dist.finalize_means(host[gi], len(g)).reshape(M, K)[:, self.top_show - 1].reshape(-1)
            out.append(MetricReport(self.metrics_list, res))
        return out

    # ------------------------------------------------------------------------------------------
    def _to_dev(self, x, dev):
        import torch
        if x is None:
            return None
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        t = x.detach().to(device=dev, dtype=torch.float32)
        if t.dim() == 2 and t.stride(1) != 1:
            t = t.contiguous()
        if t.dim() == 1:
            t = t.contiguous()
        return t

    @staticmethod
    def _host_f32(x):
        """numpy float32 view of a host array / CPU tensor with unit inner stride (no copy when possible)."""
        import torch
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            x = x.detach()
            if x.dtype != torch.float32:
                x = x.float()
            x = x.numpy()
        x = np.asarray(x, dtype=np.float32)
        if x.ndim >= 1 and x.strides[-1] != 4:
            x = np.ascontiguousarray(x)
        return x

    def _evaluate_fused(self, model, users, key, dev, want_pu, sums, host_fast, rank=0, world=1):
        """Column sums are ADDED into `sums` (float64 device tensor [M*K]) -- unless host tables took the one-call
        native path (`host_fast`), which returns them on the host.
        -> (path, float64 host sums or None, per-user block on the device or None)"""
        import torch
        with _nvtx("skrec:model.eval_embeddings"):
            user_vecs, item_vecs, bias = model.eval_embeddings(users)
        MK = self.metrics_num * self.max_top
        score_fn = getattr(model, "score_fn", "dot")
        assert score_fn in ("dot", "neg_l2"), "score_fn must be 'dot' or 'neg_l2'"
        neg_l2 = score_fn == "neg_l2"
        prec = self.precision
        if neg_l2:
            assert prec in ("auto", "fp32"), "score_fn='neg_l2' (-||u - i|| + b) runs on the FP32 tile kernels: precision auto|fp32"
            prec = "fp32"
        on_host = not (isinstance(user_vecs, torch.Tensor) and user_vecs.is_cuda) and \
            not (isinstance(item_vecs, torch.Tensor) and item_vecs.is_cuda)
        if on_host and host_fast and not want_pu:
            # host tables (numpy / CPU tensors, pinned or not): one native call does H2D, the whole
            # pipeline and the D2H of the sums, with a single synchronisation at the end
            uv, iv, b = self._host_f32(user_vecs), self._host_f32(item_vecs), self._host_f32(bias)
            assert uv.ndim == 2 and iv.ndim == 2 and uv.shape[1] == iv.shape[1], "eval_embeddings: shapes"
            assert uv.shape[0] == len(users), "eval_embeddings must return one row per requested user"
            plan = self._plan(users, int(iv.shape[0]), key)
            plan.ctx.set_option("score_fn", 1 if neg_l2 else 0)
            with _nvtx("skrec:fused (host tables: H2D + kernels + D2H)"):
                _, _, host_sums = plan.ctx.eval_fused_host(uv, iv, b, 0, self.metrics, self.max_top, precision=prec)
            return "fused:" + plan.ctx.last_fused_kernel, host_sums, None
        with _nvtx("skrec:H2D embedding tables"):
            items_on_host = not (isinstance(item_vecs, torch.Tensor) and item_vecs.is_cuda)
            if items_on_host and world > 1 and self.upload == "sharded" and getattr(item_vecs, "ndim", 0) == 2:
                # user-sharded ranks all hold the same item table on the host: each uploads 1/world of its rows and
                # the slices are all-gathered over NVLink, while the PCIe links carry this rank's user rows
                from . import dist
                iv, work = dist.gather_host_table(item_vecs, dev, rank, world, self.process_group)
                uv, b = self._to_dev(user_vecs, dev), self._to_dev(bias, dev)
                if work is not None:
                    work.wait()
            else:
                uv, iv, b = self._to_dev(user_vecs, dev), self._to_dev(item_vecs, dev), self._to_dev(bias, dev)
        if len(users) == 0:
            return "none", None, None
        assert uv.dim() == 2 and iv.dim() == 2 and uv.shape[1] == iv.shape[1], "eval_embeddings: shapes"
        assert uv.shape[0] == len(users), "eval_embeddings must return one row per requested user"
        d = int(iv.shape[1])
        if d % 4 != 0 and (d > 128 or self.max_top > 128 or prec == "fp32"):
            # the FP32 tile kernels (wide d, top-K > 128, precision="fp32") read float4: zero-pad the rows to a
            # multiple of four columns -- zeros add nothing to a dot product (the host entry pads the same way)
            uv = torch.nn.functional.pad(uv, (0, 4 - d % 4))
            iv = torch.nn.functional.pad(iv, (0, 4 - d % 4))
        plan = self._plan(users, int(iv.shape[0]), key)
        per_user = torch.empty((len(users), MK), dtype=torch.float32, device=dev) if want_pu else None
        plan.ctx.set_option("score_fn", 1 if neg_l2 else 0)
        with _nvtx("skrec:fused (split + SAMPLE + COLLECT + select + metrics)"):
            plan.ctx.eval_fused(uv, iv, b, 0, self.metrics, self.max_top, precision=prec, per_user=per_user, sums=sums)
        return "fused:" + plan.ctx.last_fused_kernel, None, per_user

    def _evaluate_item_sharded(self, model, users, key, dev, rank, world, sums):
        """Item-sharded evaluation.  Per user chunk: local top-K over this rank's item rows
        (fused kernels) -> all-gather of the [n, K] rank keys -> this rank merges and evaluates its slice
        of the chunk's users; column sums are ADDED into `sums`.  The all-gather of chunk n is issued from a side
        stream and overlaps the top-K of chunk n + 1 (double-buffered lists); every library call stays on the
        current stream.  -> (path, number of users in my slices)"""
        import inspect
        import torch
        import torch.distributed as td
        from . import dist
        try:
            takes_shard = "item_shard" in inspect.signature(model.eval_embeddings).parameters
        except (TypeError, ValueError):
            takes_shard = False
        if takes_shard:
            user_vecs, item_rows, bias_rows, n_items = model.eval_embeddings(users, item_shard=(rank, world))
            ilo, ihi = dist.shard_range(int(n_items), rank, world)
        else:
            user_vecs, item_vecs, bias = model.eval_embeddings(users)
            n_items = int(item_vecs.shape[0])
            ilo, ihi = dist.shard_range(n_items, rank, world)
            item_rows = item_vecs[ilo:ihi]
            bias_rows = None if bias is None else bias[ilo:ihi]
        uv, iv, b = self._to_dev(user_vecs, dev), self._to_dev(item_rows, dev), self._to_dev(bias_rows, dev)
        assert iv.shape[0] == ihi - ilo, "eval_embeddings(item_shard=...) must return exactly this rank's item rows"
        assert uv.shape[0] == len(users) and uv.shape[1] == iv.shape[1], "eval_embeddings: shapes"
        K = self.max_top
        assert world * K <= 1024, "shard='items': world_size * max(top_k) must not exceed 1024"
        plan = self._plan(users, int(n_items), key, (ilo, ihi))
        n_mine = 0
        chunk_rows = max(128, (int(self.item_chunk_rows) // 128) * 128)  # row offsets of the fused kernels are multiples of 128
        chunks = [(c0, min(chunk_rows, len(users) - c0)) for c0 in range(0, len(users), chunk_rows)]
        n_max = chunks[0][1]
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev)
        keys = [torch.empty((n_max, K), dtype=torch.int64, device=dev) for _ in range(min(2, len(chunks)))]
        gathered = [torch.empty((world, n_max, K), dtype=torch.int64, device=dev) for _ in range(min(2, len(chunks)))]
        self.last_gather_bytes = 0

        def merge(j, work):
            nonlocal n_mine
            c0, n = chunks[j]
            work.wait()  # the current stream waits for the all-gather of chunk j
            lo, hi = dist.shard_range(n, rank, world)
            if hi > lo:
                with _nvtx("skrec:merge + metrics of my slice"):
                    plan.ctx.eval_merged_topk(gathered[j % 2][:, :n] if n == n_max else gathered[j % 2][:, :n].contiguous(),
                                              lo, hi - lo, c0 + lo, self.metrics, K, sums=sums)
                n_mine += hi - lo

        pending = None
        for j, (c0, n) in enumerate(chunks):
            kj = keys[j % 2][:n]
            with _nvtx("skrec:topk_fused (my item rows)"):
                plan.ctx.topk_fused(uv[c0:c0 + n], iv, b, c0, ilo, K, kj, precision=self.precision)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(side), _nvtx("skrec:all-gather rank keys"):
                side.wait_event(done)
                if n == n_max:
                    out = gathered[j % 2]
                    work = td.all_gather_into_tensor(out, kj, group=self.process_group, async_op=True)
                else:  # short last chunk: gather into a dense [world, n, K] view of the buffer
                    out = gathered[j % 2].view(-1)[:world * n * K].view(world, n, K)
                    work = td.all_gather_into_tensor(out, kj, group=self.process_group, async_op=True)
                    gathered[j % 2] = out
            self.last_gather_bytes += int(world * n * K * 8)
            if pending is not None:
                merge(*pending)
            pending = (j, work)
        merge(*pending)
        return "items:" + plan.ctx.last_fused_kernel, n_mine

    def _side_stream(self, dev):
        import torch
        st = getattr(self, "_side", None)
        if st is None or st.device != dev:
            st = self._side = torch.cuda.Stream(device=dev)
        return st

    def _evaluate_predict(self, model, users, key, dev, want_pu, sums):
        """The reference protocol: `predict(batch)` blocks (evaluator.py:188-202), column sums ADDED into `sums`.
        Host blocks go through two pinned staging buffers and a copy stream: the upload of block n + 1 overlaps the
        masking / top-K / metric kernels of block n (the model's own `predict` of block n + 1 overlaps both)."""
        import torch
        plan = None
        MK = self.metrics_num * self.max_top
        per_user = torch.empty((len(users), MK), dtype=torch.float32, device=dev) if want_pu else None
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev)
        stage = getattr(self, "_stage", None)
        slot = 0
        for b0 in range(0, len(users), self.batch_size):  # sequential, last batch short (batch_iterator.py:98-106)
            batch_users = users[b0:b0 + self.batch_size]
            with _nvtx("skrec:model.predict"):
                ranking_score = model.predict(batch_users)  # (B,N)
            if isinstance(ranking_score, torch.Tensor):
                s = ranking_score.detach().to(device=dev, dtype=torch.float32)
                if s.stride(1) != 1:
                    s = s.contiguous()
            else:
                assert isinstance(ranking_score, np.ndarray), "'ranking_score' must be an np.ndarray"
                h = np.ascontiguousarray(ranking_score, dtype=np.float32)
                if stage is None or stage[0][0].numel() < h.size:
                    n_el = max(h.size, int(self.batch_size) * h.shape[1])
                    stage = self._stage = [(torch.empty(n_el, dtype=torch.float32).pin_memory(),
                                            torch.empty(n_el, dtype=torch.float32, device=dev), torch.cuda.Event()) for _ in range(2)]
                pin, dbuf, free = stage[slot]
                slot ^= 1
                free.synchronize()  # the kernels that read this slot two blocks ago are done (host may overwrite `pin`)
                pin[:h.size].view(h.shape).numpy()[...] = h
                with torch.cuda.stream(side), _nvtx("skrec:H2D score block"):
                    s = dbuf[:h.size].view(h.shape)
                    s.copy_(pin[:h.size].view(h.shape), non_blocking=True)
                    up = torch.cuda.Event()
                    up.record(side)
                main.wait_event(up)
            if plan is None:
                plan = self._plan(users, int(s.shape[1]), key)
            with _nvtx("skrec:eval_scores (mask + top-K + metrics)"):
                plan.ctx.eval_scores(s, b0, self.metrics, self.max_top,
                                     per_user=None if per_user is None else per_user[b0:b0 + len(batch_users)],
                                     sums=sums)
            if not isinstance(ranking_score, torch.Tensor):
                free.record(main)
        return "scores", per_user
