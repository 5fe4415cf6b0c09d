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

_metric2id = {"Precision": 1, "Recall": 2, "MAP": 3, "NDCG": 4, "MRR": 5}
_id2metric = {value: key for key, value in _metric2id.items()}


def _dict_to_csr(users, d, col_range=None):
    """{user: int array} restricted to `users` (in order) -> (indptr int64, indices int32).
    col_range=(lo, hi): keep only items in [lo, hi) and renumber them from 0 (an item shard's column
    partition of the interactions)."""
    counts = np.fromiter((len(d[u]) if u in d else 0 for u in users), dtype=np.int64, count=len(users))
    indptr = np.zeros(len(users) + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    if indptr[-1] == 0:
        return indptr, np.zeros(0, np.int32)
    indices = np.concatenate([np.asarray(d[u], dtype=np.int32).ravel() for u in users if u in d and len(d[u]) > 0])
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
        from . import _native
        self.users = users
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


class RankingEvaluator(object):
    """Evaluator for item ranking task (reference evaluator.py:61-214), GPU-resident.

    Args (reference): user_train_dict, user_test_dict, metric, top_k, batch_size, num_thread.
        `num_thread` is accepted and stored (bert4rec_utils.py:79 reads it) but unused: the CUDA
        grid replaces the thread pool.  `batch_size` is the user batch of the `predict` path.
    Keyword-only additions:
        device: CUDA device index (default: current torch device).
        precision: "auto" | "3xtf32" | "fp32" | "tf32r" | "1xtf32" -- arithmetic of the fused scoring
            ("tf32r": one TF32 pass finds candidates inside a rigorous error band, the survivors are
            re-scored in exact FP32; same results as "fp32").
        mean: "f64" (float64 sums, rounded once to float32) or "numpy_f32" (the reference's
            float32 row-order accumulation of np.mean, evaluator.py:208, bit for bit;
            single-process only).
        shard_users: with torch.distributed initialised, each rank evaluates a contiguous slice
            of the users and the metric sums are all-reduced (default True).
        shard: "users" (default; the item table is replicated) or "items" (catalogue beyond one HBM:
            every rank holds a contiguous range of item rows, computes every user's top-K over its
            range, the per-rank lists are all-gathered and merged, then the sums all-reduced).  With
            "items" the model's `eval_embeddings` may take `item_shard=(rank, world)` and return just
            its rows plus the catalogue size, `(user_vecs, item_rows, bias_rows | None, n_items)`.
        process_group: the group to reduce over (default: WORLD).
    """

    def __init__(self, user_train_dict: Optional[Dict[int, np.ndarray]],
                 user_test_dict: Dict[int, np.ndarray],
                 metric: Union[None, str, Tuple[str], List[str]] = None,
                 top_k: Union[int, List[int], Tuple[int]] = 50,
                 batch_size: int = 256, num_thread: int = 8, *,
                 device: Optional[int] = None, precision: str = "auto", mean: str = "f64",
                 shard_users: bool = True, shard: str = "users", process_group=None):
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

        assert precision in ("auto", "3xtf32", "fp32", "1xtf32", "tf32r"), "precision must be auto|3xtf32|fp32|1xtf32|tf32r"
        assert mean in ("f64", "numpy_f32"), "mean must be f64|numpy_f32"
        assert shard in ("users", "items"), "shard must be users|items"
        self.shard = shard
        self.device = device
        self.precision = precision
        self.mean = mean
        self.shard_users = shard_users
        self.process_group = process_group
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
        self.user_pos_train = user_train_dict if user_train_dict is not None else dict()
        self._plans = OrderedDict()

    def set_test_data(self, user_test_dict: Dict[int, np.ndarray]):
        assert len(user_test_dict) > 0, "'user_test_dict' can be empty."
        self.user_pos_test = user_test_dict
        self._all_users = list(user_test_dict.keys())  # evaluation order of evaluate(model) (evaluator.py:184)
        self._plans = OrderedDict()

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

    def _plan(self, users, n_items, key, item_range=None):
        plan = self._plans.get(key)
        if plan is not None and plan.n_items == n_items and plan.item_range == item_range:
            self._plans.move_to_end(key)
            return plan
        plan = _Plan(users, n_items, self.user_pos_train, self.user_pos_test, self._device_index(), item_range)
        self._plans[key] = plan
        while len(self._plans) > 6:
            self._plans.popitem(last=False)
        return plan

    def _shard(self, n):
        """(rank, world, lo, hi): this rank's contiguous slice of n evaluated users."""
        from . import dist
        if not self.shard_users:
            return 0, 1, 0, n
        rank, world = dist.rank_world(self.process_group)
        lo, hi = dist.shard_range(n, rank, world)
        return rank, world, lo, hi

    def evaluate(self, model, test_users: Optional[Iterable[int]] = None) -> MetricReport:
        """Evaluate `model` (reference evaluator.py:163-214).

        `model` must have `predict(users) -> float32 ndarray [B, num_items]`; if it also has
        `eval_embeddings(users)` the fused path is used.
        """
        import torch
        from . import dist

        assert hasattr(model, "predict") or hasattr(model, "eval_embeddings"), "the model must have attribute 'predict'."
        if test_users is not None:
            test_users = [u for u in test_users if u in self.user_pos_test]
            key_all = ("subset", hash(tuple(test_users)), len(test_users))
        else:
            if len(self._all_users) != len(self.user_pos_test):  # the dict was mutated behind our back
                self._all_users = list(self.user_pos_test.keys())
                self._plans = OrderedDict()
            test_users = self._all_users
            key_all = ("all",)
        assert isinstance(test_users, Iterable), "'test_user' must be iterable."

        rank, world, lo, hi = self._shard(len(test_users))
        item_sharded = self.shard == "items" and world > 1
        if item_sharded:
            assert hasattr(model, "eval_embeddings"), "shard='items' needs a model with eval_embeddings"
            assert self.mean == "f64", "shard='items' supports mean='f64' only"
            lo, hi = 0, len(test_users)  # every rank sees every user; the items are what is split
        users = test_users[lo:hi] if world > 1 else test_users
        key = key_all + (rank, world, self.shard)
        dev = torch.device("cuda", self._device_index())
        K, M = self.max_top, self.metrics_num
        MK = M * K
        per_user = None
        path = "none"
        n_mine = 0
        col_sums = np.zeros(MK, dtype=np.float64)  # this rank's float64 column sums
        want_pu = self.mean == "numpy_f32"

        if len(users) > 0:
            with torch.cuda.device(dev):
                if item_sharded:
                    path, col_sums, n_mine = self._evaluate_item_sharded(model, users, key, dev, rank, world)
                elif hasattr(model, "eval_embeddings"):
                    path, col_sums, per_user = self._evaluate_fused(model, users, key, dev, want_pu)
                else:
                    path, col_sums, per_user = self._evaluate_predict(model, users, key, dev, want_pu)

        if self.mean == "numpy_f32":
            if world > 1:
                raise RuntimeError("mean='numpy_f32' reproduces a sequential sum and is single-process only")
            plan = self._plans[key]
            acc = torch.zeros(MK, dtype=torch.float32, device=dev)
            plan.ctx.colsum_f32_seq(per_user, acc)
            final_results = (acc / torch.tensor(float(len(users)), dtype=torch.float32, device=dev)).cpu().numpy()
        else:
            n_users = float(n_mine) if item_sharded else float(len(users))
            if world > 1:
                packed = torch.from_numpy(np.concatenate([col_sums, [n_users]])).to(dev)  # [column sums | user count]
                dist.allreduce_sums(packed, self.process_group)
                host = packed.cpu().numpy()
                col_sums, n_users = host[:MK], host[MK]
            final_results = dist.finalize_means(col_sums, n_users)

        self.last_stats = {"path": path, "users": len(users), "world": world}
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
        key = ("subset", hash(tuple(union)), len(union), 0, 1, "groups")
        with torch.cuda.device(dev):
            if hasattr(model, "eval_embeddings"):
                _, _, per_user = self._evaluate_fused(model, union, key, dev, True)
            else:
                _, _, per_user = self._evaluate_predict(model, union, key, dev, True)
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

    def _fused_can_take(self, d, n_items=None):
        """Shapes the fused kernels cover: top-K <= 128; d <= 128 on tensor cores, else FP32 FMA with d % 4 == 0.
        With precision "auto" a catalogue too small for sampled thresholds (about 160 items per requested rank:
        ml-1m at top-50) goes to the score-block path as well -- the fused kernels would still be exact there, but
        through their per-row fallback, which is slower than a GEMM plus the HBM-bound top-K kernel."""
        if self.max_top > 128:
            return False
        if self.precision in ("3xtf32", "tf32r", "1xtf32"):
            return True  # an explicit tensor-core request fails loudly in the library if the shape is out of range
        if self.precision == "auto" and n_items is not None and n_items < max(3072, 160 * self.max_top):
            return False
        return d <= 128 or (d % 4 == 0 and d <= 1024)

    def _evaluate_by_blocks(self, uv, iv, b, users, key, dev, want_pu):
        """Shapes outside the fused kernels (top-K > 128, odd wide d): score blocks of `batch_size` users on the
        device (`U_b @ I^T + b`, FP32 -- the one place a library GEMM is used) and feed them to the
        score-matrix kernels, like the `predict` path but without the host round trip."""
        import torch
        MK = self.metrics_num * self.max_top
        plan = self._plan(users, int(iv.shape[0]), key)
        sums = torch.zeros(MK, dtype=torch.float64, device=dev)
        per_user = torch.empty((len(users), MK), dtype=torch.float32, device=dev) if want_pu else None
        step = max(1, min(int(self.batch_size), 4096))
        for b0 in range(0, len(users), step):
            s = torch.matmul(uv[b0:b0 + step].float(), iv.float().T)
            if b is not None:
                s = s + b
            plan.ctx.eval_scores(s.contiguous(), b0, self.metrics, self.max_top,
                                 per_user=None if per_user is None else per_user[b0:b0 + s.shape[0]], sums=sums)
        return "scores:device_blocks", sums.cpu().numpy(), per_user

    def _evaluate_fused(self, model, users, key, dev, want_pu):
        """-> (path, float64 column sums [M*K] on the host, per-user block on the device or None)"""
        import torch
        user_vecs, item_vecs, bias = model.eval_embeddings(users)
        MK = self.metrics_num * self.max_top
        if not self._fused_can_take(int(item_vecs.shape[1]), int(item_vecs.shape[0])):
            uv, iv, b = self._to_dev(user_vecs, dev), self._to_dev(item_vecs, dev), self._to_dev(bias, dev)
            assert uv.shape[0] == len(users) and uv.shape[1] == iv.shape[1], "eval_embeddings: shapes"
            return self._evaluate_by_blocks(uv, iv, b, users, key, dev, want_pu)
        on_host = not (isinstance(user_vecs, torch.Tensor) and user_vecs.is_cuda) and \
            not (isinstance(item_vecs, torch.Tensor) and item_vecs.is_cuda)
        if on_host and not want_pu:
            # host tables (numpy / CPU tensors, pinned or not): one native call does H2D, the whole
            # pipeline and the D2H of the sums, with a single synchronisation at the end
            uv, iv, b = self._host_f32(user_vecs), self._host_f32(item_vecs), self._host_f32(bias)
            assert uv.ndim == 2 and iv.ndim == 2 and uv.shape[1] == iv.shape[1], "eval_embeddings: shapes"
            assert uv.shape[0] == len(users), "eval_embeddings must return one row per requested user"
            plan = self._plan(users, int(iv.shape[0]), key)
            _, _, sums = plan.ctx.eval_fused_host(uv, iv, b, 0, self.metrics, self.max_top, precision=self.precision)
            return "fused:" + plan.ctx.last_fused_kernel, sums, None
        uv, iv, b = self._to_dev(user_vecs, dev), self._to_dev(item_vecs, dev), self._to_dev(bias, dev)
        assert uv.dim() == 2 and iv.dim() == 2 and uv.shape[1] == iv.shape[1], "eval_embeddings: shapes"
        assert uv.shape[0] == len(users), "eval_embeddings must return one row per requested user"
        plan = self._plan(users, int(iv.shape[0]), key)
        sums = torch.zeros(MK, dtype=torch.float64, device=dev)
        per_user = torch.empty((len(users), MK), dtype=torch.float32, device=dev) if want_pu else None
        plan.ctx.eval_fused(uv, iv, b, 0, self.metrics, self.max_top, precision=self.precision, per_user=per_user, sums=sums)
        return "fused:" + plan.ctx.last_fused_kernel, sums.cpu().numpy(), per_user

    def _evaluate_item_sharded(self, model, users, key, dev, rank, world, chunk_rows=1 << 18):
        """Item-sharded evaluation.  Per user chunk: local top-K over this rank's item rows
        (fused kernels) -> all-gather of the [n, K] rank keys -> this rank merges and evaluates its slice
        of the chunk's users.  -> (path, float64 column sums of my slices, number of users in them)"""
        import inspect
        import torch
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
        K, MK = self.max_top, self.metrics_num * self.max_top
        assert world * K <= 1024, "shard='items': world_size * max(top_k) must not exceed 1024"
        plan = self._plan(users, int(n_items), key, (ilo, ihi))
        sums = torch.zeros(MK, dtype=torch.float64, device=dev)
        n_mine = 0
        chunk_rows = max(128, (int(chunk_rows) // 128) * 128)  # row offsets of the fused kernels are multiples of 128
        for c0 in range(0, len(users), chunk_rows):
            n = min(chunk_rows, len(users) - c0)
            keys = torch.empty((n, K), dtype=torch.int64, device=dev)
            plan.ctx.topk_fused(uv[c0:c0 + n], iv, b, c0, ilo, K, keys, precision=self.precision)
            keys_all = dist.allgather_keys(keys, self.process_group)  # [world, n, K]
            lo, hi = dist.shard_range(n, rank, world)
            if hi > lo:
                plan.ctx.eval_merged_topk(keys_all, lo, hi - lo, c0 + lo, self.metrics, K, sums=sums)
                n_mine += hi - lo
        return "items:" + plan.ctx.last_fused_kernel, sums.cpu().numpy(), n_mine

    def _evaluate_predict(self, model, users, key, dev, want_pu):
        import torch
        plan = None
        MK = self.metrics_num * self.max_top
        sums = torch.zeros(MK, dtype=torch.float64, device=dev)
        per_user = torch.empty((len(users), MK), dtype=torch.float32, device=dev) if want_pu else None
        for b0 in range(0, len(users), self.batch_size):  # sequential, last batch short (batch_iterator.py:98-106)
            batch_users = users[b0:b0 + self.batch_size]
            ranking_score = model.predict(batch_users)  # (B,N)
            if isinstance(ranking_score, torch.Tensor):
                s = ranking_score.detach().to(device=dev, dtype=torch.float32)
            else:
                assert isinstance(ranking_score, np.ndarray), "'ranking_score' must be an np.ndarray"
                s = torch.from_numpy(np.ascontiguousarray(ranking_score, dtype=np.float32)).to(dev)
            if s.stride(1) != 1:
                s = s.contiguous()
            if plan is None:
                plan = self._plan(users, int(s.shape[1]), key)
            plan.ctx.eval_scores(s, b0, self.metrics, self.max_top,
                                 per_user=None if per_user is None else per_user[b0:b0 + len(batch_users)],
                                 sums=sums)
        return "scores", sums.cpu().numpy(), per_user
