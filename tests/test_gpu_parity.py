"""Parity of the CUDA path (through the C ABI) against the oracle, the golden fixtures made by the
compiled reference, and -- where oracle/_ref travelled -- the compiled reference itself.

Bars: bit-exact for everything integer/index and for per-user float32 metric vectors given the
same score matrix; for the fused path (scores produced on chip) mean metrics within 1e-5 absolute
and rank-list differences only at near-ties |dscore| < 1e-5 (BASELINE.json north_star).
"""
import glob
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

TOL_METRIC = 1e-5   # north_star: metrics within 1e-5 absolute
TOL_NEAR_TIE = 1e-5  # north_star: rank-list differences only where |dscore| < 1e-5
TOL_SCORE = 2e-6    # 3xTF32 / FP32-FMA score vs FP64-exact score (3e-7 measured)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback to test")
    return torch


@pytest.fixture(scope="module")
def ctx(torch_cuda):
    from skrec_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()


def _tie_free(g, B, N, scale=1.0):
    return np.stack([(g.permutation(N).astype(np.float32) - N / 2) * np.float32(scale / N) for _ in range(B)])


def _rand_csr(g, B, N, max_n, min_n=0):
    sizes = g.integers(min_n, max_n + 1, size=B)
    indptr = np.zeros(B + 1, np.int64)
    np.cumsum(sizes, out=indptr[1:])
    rows = [g.choice(N, size=int(n), replace=False) for n in sizes]
    return indptr, (np.concatenate(rows) if indptr[-1] else np.zeros(0)).astype(np.int32)


def _run_scores(torch, ctx, s, tr, te, metric, K, ld=None, row0=0):
    B, N = s.shape
    ctx.set_train_csr(tr[0], tr[1], N) if tr is not None else ctx.set_train_csr(None, None, N)
    ctx.set_test_csr(te[0], te[1], N)
    if ld is None:
        sd = torch.from_numpy(s).cuda()
    else:
        buf = torch.full((B, ld), 7.0, dtype=torch.float32, device="cuda")
        buf[:, :N] = torch.from_numpy(s).cuda()
        sd = buf[:, :N]
    MK = len(metric) * K
    idx = torch.empty((B, K), dtype=torch.int32, device="cuda")
    val = torch.empty((B, K), dtype=torch.float32, device="cuda")
    per = torch.empty((B, MK), dtype=torch.float32, device="cuda")
    sums = torch.zeros(MK, dtype=torch.float64, device="cuda")
    ctx.eval_scores(sd, row0, metric, K, topk_idx=idx, topk_val=val, per_user=per, sums=sums)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), val.cpu().numpy(), per.cpu().numpy(), sums.cpu().numpy()


def _oracle_scores_path(s, tr, te, metric, K):
    m = s.copy()
    if tr is not None:
        oracle.mask_rows(m, tr[0], tr[1])
    per, top = oracle.eval_scores(m, te[0], te[1], metric, K, return_topk=True)
    return per, top, m


@pytest.mark.parametrize("B,N,K,ld", [(5, 40, 3, None), (64, 257, 10, None), (33, 1500, 50, 1504), (17, 4099, 100, None),
                                       (9, 12345, 20, 12347), (3, 70000, 50, None), (4, 600, 512, None)])
def test_scores_path_bit_exact(torch_cuda, ctx, B, N, K, ld):
    g = np.random.default_rng(B * 1000 + N)
    s = _tie_free(g, B, N, scale=3.0)
    tr = _rand_csr(g, B, N, min(60, N // 4))
    te = _rand_csr(g, B, N, 25)
    metric = [1, 2, 3, 4, 5]
    idx, val, per, sums = _run_scores(torch_cuda, ctx, s, tr, te, metric, K, ld=ld)
    eper, etop, masked = _oracle_scores_path(s, tr, te, metric, K)
    assert np.array_equal(idx, etop)
    assert np.array_equal(val, np.take_along_axis(masked, etop.astype(np.int64), 1))
    assert np.array_equal(per, eper)
    assert np.allclose(sums, oracle.sums_f64(eper), rtol=0, atol=1e-9)


def test_scores_path_tie_policy_and_special_values(torch_cuda, ctx):
    g = np.random.default_rng(42)
    B, N, K = 40, 3000, 30
    s = g.integers(0, 12, size=(B, N)).astype(np.float32)  # heavy ties
    s[0, :50] = -np.inf
    s[1, 7] = np.nan
    s[2] = 0.0
    s[2, ::2] = -0.0  # -0.0 == +0.0 for the reference's comparator
    s[3] = -np.inf
    te = _rand_csr(g, B, N, 10)
    idx, val, per, _ = _run_scores(torch_cuda, ctx, s, None, te, [1, 4], K)
    eper, etop, _ = _oracle_scores_path(s, None, te, [1, 4], K)
    assert np.array_equal(idx, etop)
    assert np.array_equal(per, eper)
    clean = np.where(np.isnan(s), -np.inf, s)
    for r in range(B):
        assert idx[r].tolist() == np.argsort(-clean[r], kind="stable")[:K].tolist()


def test_scores_path_row_offset_and_unsorted_duplicate_csr(torch_cuda, ctx):
    g = np.random.default_rng(7)
    U, N, K = 50, 900, 20
    s = _tie_free(g, U, N)
    tr = _rand_csr(g, U, N, 40)
    te = _rand_csr(g, U, N, 12, min_n=1)
    # shuffle within rows and add duplicates: the library sorts/dedups like the reference's set
    te_idx = te[1].copy()
    for r in range(U):
        g.shuffle(te_idx[te[0][r]:te[0][r + 1]])
    ctx.set_train_csr(tr[0], tr[1], N)
    ctx.set_test_csr(te[0], te_idx, N)
    torch = torch_cuda
    lo, hi = 13, 41
    sd = torch.from_numpy(s[lo:hi]).cuda()
    per = torch.empty((hi - lo, 3 * K), dtype=torch.float32, device="cuda")
    ctx.eval_scores(sd, lo, [2, 3, 5], K, per_user=per)
    eper, _, _ = _oracle_scores_path(s, tr, te, [2, 3, 5], K)
    assert np.array_equal(per.cpu().numpy(), eper[lo:hi])


def test_golden_matrix_fixtures_through_drop_in(torch_cuda, golden_dir):
    from skrec_b200 import eval_score_matrix
    files = sorted(glob.glob(os.path.join(golden_dir, "matrix_*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        ptr = z["test_indptr"]
        items = [z["test_indices"][ptr[r]:ptr[r + 1]] for r in range(len(ptr) - 1)]
        got = eval_score_matrix(z["scores"], items, z["metric"].tolist(), int(z["top_k"]), 4)
        assert got.dtype == np.float32 and np.array_equal(got, z["expected"]), os.path.basename(f)


def test_drop_in_equals_compiled_reference(torch_cuda):
    if not oracle.ref_python_available():
        pytest.skip("oracle/_ref (Cython build of the reference) not present")
    from skrec_b200 import eval_score_matrix
    g = np.random.default_rng(99)
    s = _tie_free(g, 70, 2111)
    items = [g.choice(2111, size=int(g.integers(0, 15)), replace=False).astype(np.int32) for _ in range(70)]
    ref = oracle.ref_eval_score_matrix(s.copy(), items, [1, 2, 3, 4, 5], 40, 4)
    got = eval_score_matrix(s, items, [1, 2, 3, 4, 5], 40, 4)
    assert np.array_equal(got, ref)


def test_metrics_from_topk_bit_exact(torch_cuda, ctx):
    torch = torch_cuda
    g = np.random.default_rng(3)
    B, N, K = 200, 5000, 100
    te = _rand_csr(g, B, N, 30)
    ctx.set_test_csr(te[0], te[1], N)
    ranks = np.stack([g.choice(N, size=K, replace=False) for _ in range(B)]).astype(np.int32)
    for r in range(0, B, 3):  # make sure there are hits
        n = te[0][r + 1] - te[0][r]
        if n:
            ranks[r, : min(n, 5)] = te[1][te[0][r]: te[0][r] + min(n, 5)]
    per = torch.empty((B, 5 * K), dtype=torch.float32, device="cuda")
    ctx.metrics_from_topk(torch.from_numpy(ranks).cuda(), 0, [1, 2, 3, 4, 5], K, per_user=per)
    exp = np.zeros((B, 5 * K), np.float32)
    import ctypes
    L = oracle.lib()
    for r in range(B):
        t = np.ascontiguousarray(te[1][te[0][r]:te[0][r + 1]])
        m = np.array([1, 2, 3, 4, 5], np.int32)
        L.skr_oracle_metrics_row(ranks[r].ctypes.data_as(ctypes.c_void_p), K, t.ctypes.data_as(ctypes.c_void_p), int(t.size),
                                 m.ctypes.data_as(ctypes.c_void_p), 5, exp[r].ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(per.cpu().numpy(), exp)


def test_numpy_f32_mean_kernel(torch_cuda, ctx):
    torch = torch_cuda
    g = np.random.default_rng(1)
    a = g.random((20011, 150)).astype(np.float32)
    acc = torch.zeros(150, dtype=torch.float32, device="cuda")
    ctx.colsum_f32_seq(torch.from_numpy(a).cuda(), acc)
    got = (acc / torch.tensor(float(a.shape[0]), dtype=torch.float32, device="cuda")).cpu().numpy()
    assert np.array_equal(got, np.mean(a, axis=0))


# ------------------------------------------------------------------------------- fused path
def _fused_case(seed, U, I, d, bias, max_train):
    g = np.random.default_rng(seed)
    ue = (g.standard_normal((U, d)) * 0.1).astype(np.float32)
    ie = (g.standard_normal((I, d)) * 0.1).astype(np.float32)
    b = (g.standard_normal(I) * 0.01).astype(np.float32) if bias else None
    tr = _rand_csr(g, U, I, max_train)
    te = _rand_csr(g, U, I, 20, min_n=1)
    return ue, ie, b, tr, te


def _run_fused(torch, ctx, ue, ie, b, tr, te, metric, K, precision, chunks=0):
    U, I = ue.shape[0], ie.shape[0]
    ctx.set_train_csr(tr[0], tr[1], I) if tr is not None else ctx.set_train_csr(None, None, I)
    ctx.set_test_csr(te[0], te[1], I)
    ctx.set_option("chunks", chunks)
    MK = len(metric) * K
    idx = torch.empty((U, K), dtype=torch.int32, device="cuda")
    val = torch.empty((U, K), dtype=torch.float32, device="cuda")
    per = torch.empty((U, MK), dtype=torch.float32, device="cuda")
    sums = torch.zeros(MK, dtype=torch.float64, device="cuda")
    ctx.eval_fused(torch.from_numpy(ue).cuda(), torch.from_numpy(ie).cuda(), None if b is None else torch.from_numpy(b).cuda(),
                   0, metric, K, precision=precision, topk_idx=idx, topk_val=val, per_user=per, sums=sums)
    torch.cuda.synchronize()
    if precision in ("tf32r", "f16r"):
        # metrics without lists: exact scores only for the test items among the survivors and their near-ties (k_select.cuh
        # HITS) -- the per-user block must come out bit for bit as from the full re-scoring above
        per2 = torch.full((U, MK), -1.0, dtype=torch.float32, device="cuda")
        sums2 = torch.zeros(MK, dtype=torch.float64, device="cuda")
        ctx.eval_fused(torch.from_numpy(ue).cuda(), torch.from_numpy(ie).cuda(), None if b is None else torch.from_numpy(b).cuda(),
                       0, metric, K, precision=precision, per_user=per2, sums=sums2)
        torch.cuda.synchronize()
        assert torch.equal(per2, per), "hits-only re-scoring changed a per-user metric"
        assert float((sums2 - sums).abs().max()) < 1e-9
        sums3 = torch.zeros(MK, dtype=torch.float64, device="cuda")  # sums only (what evaluate() asks for)
        ctx.eval_fused(torch.from_numpy(ue).cuda(), torch.from_numpy(ie).cuda(), None if b is None else torch.from_numpy(b).cuda(),
                       0, metric, K, precision=precision, sums=sums3)
        torch.cuda.synchronize()
        assert float((sums3 - sums).abs().max()) < 1e-9
    ctx.set_option("chunks", 0)
    return idx.cpu().numpy(), val.cpu().numpy(), per.cpu().numpy(), sums.cpu().numpy()


def _check_fused(idx, val, per, sums, ue, ie, b, tr, te, metric, K, tol_score=TOL_SCORE):
    U = ue.shape[0]
    S = oracle.scores(ue, ie, b)
    if tr is not None:
        oracle.mask_rows(S, tr[0], tr[1])
    eper, etop = oracle.eval_scores(S, te[0], te[1], metric, K, return_topk=True)
    got_s = np.take_along_axis(S, idx.astype(np.int64), 1)
    exp_s = np.take_along_axis(S, etop.astype(np.int64), 1)
    finite = np.isfinite(got_s)
    # (a) the scores the GPU reports are the exact scores of the items it reports
    assert np.max(np.abs(val[finite] - got_s[finite])) <= tol_score
    # (b) rank lists differ only at near-ties
    diff = idx != etop
    if diff.any():
        assert np.max(np.abs(got_s[diff] - exp_s[diff])) < TOL_NEAR_TIE
    assert diff.mean() < 0.02
    # every row is a permutation-free list
    assert all(len(set(idx[r].tolist())) == K for r in range(U))
    # (c) metrics
    mean_got = sums / U
    mean_exp = oracle.sums_f64(eper) / U
    assert np.max(np.abs(mean_got - mean_exp)) <= TOL_METRIC
    same = ~diff.any(axis=1)
    assert np.array_equal(per[same], eper[same])  # identical rank list -> identical float32 metric vector
    return diff.sum()


FUSED_SHAPES = [
    # U,    I,    d,  bias,  K, max_train
    (300, 1000, 64, True, 10, 40),
    (129, 777, 64, False, 50, 30),
    (1000, 5000, 32, True, 20, 100),
    (257, 4097, 128, True, 100, 64),
    (64, 300, 96, False, 5, 10),
]


@pytest.mark.parametrize("U,I,d,bias,K,max_train", FUSED_SHAPES)
def test_fused_fp32_matches_oracle(torch_cuda, ctx, U, I, d, bias, K, max_train):
    ue, ie, b, tr, te = _fused_case(U + I, U, I, d, bias, max_train)
    out = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, [1, 2, 3, 4, 5], K, "fp32")
    assert ctx.last_fused_kernel == "simt_fp32"
    _check_fused(*out, ue, ie, b, tr, te, [1, 2, 3, 4, 5], K)


@pytest.mark.parametrize("U,I,d,bias,K,max_train", FUSED_SHAPES)
def test_fused_3xtf32_matches_oracle(torch_cuda, ctx, U, I, d, bias, K, max_train):
    ue, ie, b, tr, te = _fused_case(U + I, U, I, d, bias, max_train)
    out = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, [1, 2, 3, 4, 5], K, "3xtf32")
    assert ctx.last_fused_kernel == "tcgen05_3xtf32"
    _check_fused(*out, ue, ie, b, tr, te, [1, 2, 3, 4, 5], K)


@pytest.mark.parametrize("precision", ["fp32", "3xtf32"])
@pytest.mark.parametrize("chunks", [1, 3, 7])
def test_fused_item_chunking_is_invisible(torch_cuda, ctx, precision, chunks):
    ue, ie, b, tr, te = _fused_case(5, 200, 3000, 64, True, 50)
    out = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, [1, 2, 4], 50, precision, chunks=chunks)
    _check_fused(*out, ue, ie, b, tr, te, [1, 2, 4], 50)


def test_fused_fewer_unmasked_items_than_k(torch_cuda, ctx):
    # I - deg(u) < K: the reference lets masked (-inf) items into the list;
    # here they follow in ascending id order.
    g = np.random.default_rng(0)
    U, I, d, K = 130, 140, 64, 100
    ue = (g.standard_normal((U, d)) * 0.1).astype(np.float32)
    ie = (g.standard_normal((I, d)) * 0.1).astype(np.float32)
    tr = _rand_csr(g, U, I, 120, min_n=60)
    te = _rand_csr(g, U, I, 5, min_n=1)
    for prec in ("fp32", "3xtf32"):
        idx, val, per, sums = _run_fused(torch_cuda, ctx, ue, ie, None, tr, te, [1, 2], K, prec)
        S = oracle.scores(ue, ie, None)
        oracle.mask_rows(S, tr[0], tr[1])
        _, etop = oracle.eval_scores(S, te[0], te[1], [1, 2], K, return_topk=True)
        n_unmasked = I - np.diff(tr[0])
        for r in range(U):
            n = int(min(K, n_unmasked[r]))
            assert set(idx[r, :n].tolist()) == set(etop[r, :n].tolist())
            assert idx[r, n:].tolist() == etop[r, n:].tolist()  # masked tail, ascending id


def test_fused_host_entry_and_errors(torch_cuda, ctx):
    from skrec_b200 import _native
    ue, ie, b, tr, te = _fused_case(11, 150, 900, 20, True, 30)  # d=20: padded inside the call
    ctx.set_train_csr(tr[0], tr[1], 900)
    ctx.set_test_csr(te[0], te[1], 900)
    per, idx, sums = ctx.eval_fused_host(ue, ie, b, 0, [1, 2, 4], 10, precision="auto", want_topk=True, want_per_user=True)
    val = np.take_along_axis(oracle.scores(ue, ie, b), idx.astype(np.int64), 1)
    _check_fused(idx, val, per, sums, ue, ie, b, tr, te, [1, 2, 4], 10)
    with pytest.raises(_native.NativeError):  # evaluate.h:45 reads out of bounds when N < K; refused here
        ctx.eval_fused_host(ue, ie[:5], None, 0, [1], 10)
    with pytest.raises(_native.NativeError):
        ctx.eval_fused_host(ue, ie, b, 0, [9], 10)


def test_single_tf32_pass_is_not_reference_grade_but_runs(torch_cuda, ctx):
    ue, ie, b, tr, te = _fused_case(21, 256, 4000, 64, False, 20)
    idx, val, per, sums = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, [4], 20, "1xtf32")
    assert ctx.last_fused_kernel == "tcgen05_1xtf32"
    S = oracle.scores(ue, ie, b)
    err = np.max(np.abs(val - np.take_along_axis(S, idx.astype(np.int64), 1)))
    assert 1e-6 < err < 5e-3  # visibly worse than 3xTF32, still a dot product


# ------------------------------------------------------------------------------- evaluator class
class _PredictModel(object):
    def __init__(self, ue, ie, b, users):
        self.ue, self.ie, self.b = ue, ie, b

    def predict(self, users):
        s = self.ue[np.asarray(users)] @ self.ie.T
        if self.b is not None:
            s = s + self.b
        return np.ascontiguousarray(s, dtype=np.float32)


class _FusedModel(_PredictModel):
    def eval_embeddings(self, users):
        return self.ue[np.asarray(users)], self.ie, self.b


def _golden_evaluator(z):
    users = z["users"].tolist()
    trp, tri, tep, tei = z["train_indptr"], z["train_indices"], z["test_indptr"], z["test_indices"]
    train = {u: tri[trp[i]:trp[i + 1]] for i, u in enumerate(users) if trp[i + 1] > trp[i]}
    test = {u: tei[tep[i]:tep[i + 1]] for i, u in enumerate(users)}
    names = {1: "Precision", 2: "Recall", 3: "MAP", 4: "NDCG", 5: "MRR"}
    return train, test, [names[int(m)] for m in z["metric"]], z["top_k"].tolist(), (z["bias"] if z["bias"].size else None)


def test_evaluator_class_against_reference_fixtures(torch_cuda, golden_dir):
    import re
    from skrec_b200 import RankingEvaluator
    for f in sorted(glob.glob(os.path.join(golden_dir, "evaluator_*.npz"))):
        z = np.load(f)
        train, test, metric, top_k, bias = _golden_evaluator(z)
        # score-matrix path, numpy-order float32 mean: the reference's digits, bit for bit
        ev = RankingEvaluator(train, test, metric=metric, top_k=top_k, batch_size=41, mean="numpy_f32")
        rep = ev.evaluate(_PredictModel(z["user_emb"], z["item_emb"], bias, None))
        assert list(rep.metrics()) == z["expected_names"].tolist()
        assert np.array_equal(np.array(list(rep.values()), np.float32), z["expected_values"]), os.path.basename(f)
        assert re.sub(r"\x1b\[[0-9]+m", "", rep.values_str) == str(z["values_str"])
        assert re.sub(r"\x1b\[[0-9]+m", "", ev.metrics_str) == str(z["metrics_str"])
        assert ev.last_stats["path"] == "scores"
        # default float64 sums: within float32 rounding of the same numbers
        rep64 = RankingEvaluator(train, test, metric=metric, top_k=top_k).evaluate(_PredictModel(z["user_emb"], z["item_emb"], bias, None))
        assert np.max(np.abs(np.array(list(rep64.values())) - z["expected_values"])) < 1e-6
        # fused paths
        for prec in ("fp32", "3xtf32"):
            evf = RankingEvaluator(train, test, metric=metric, top_k=top_k, precision=prec)
            repf = evf.evaluate(_FusedModel(z["user_emb"], z["item_emb"], bias, None))
            assert evf.last_stats["path"].startswith("fused:")
            assert np.max(np.abs(np.array(list(repf.values())) - z["expected_values"])) <= TOL_METRIC, (os.path.basename(f), prec)


def test_evaluator_test_users_subset_and_groups(torch_cuda, golden_dir):
    from skrec_b200 import RankingEvaluator
    z = np.load(os.path.join(golden_dir, "evaluator_bias.npz"))
    train, test, metric, top_k, bias = _golden_evaluator(z)
    model = _PredictModel(z["user_emb"], z["item_emb"], bias, None)
    ev = RankingEvaluator(train, test, metric=metric, top_k=top_k, mean="numpy_f32")
    subset = [u for u in list(test.keys())[::3]] + [10 ** 6]  # unknown users are filtered (evaluator.py:182)
    rep = ev.evaluate(model, test_users=subset)
    per, _ = oracle.evaluate_dicts(model.predict, train, test, [int(m) for m in z["metric"]], max(top_k), users=subset)
    K = max(top_k)
    exp = oracle.mean_f32(per).reshape(len(metric), K)[:, np.sort(top_k) - 1].ravel()
    assert np.array_equal(np.array(list(rep.values()), np.float32), exp)
    if oracle.ref_python_available():
        ref = oracle.RefRankingEvaluator(train, test, metric=metric, top_k=top_k).evaluate(model, test_users=subset)
        assert np.array_equal(np.array(list(ref.values()), np.float32), exp)


# ------------------------------------------------------------------------------- full-size properties
@pytest.fixture(scope="module")
def c2_data(torch_cuda):
    from skrec_b200 import synth
    return synth.make_config("c2", device="cuda")


def test_c2_full_size_fused_vs_oracle_sample_and_cross_kernel(torch_cuda, ctx, c2_data):
    """BASELINE.json configs[1] at full size: the two fused kernels agree with each other on every
    user (metrics 1e-5, lists differ only at near-ties) and with the oracle on a 512-user sample."""
    torch = torch_cuda
    d = c2_data
    tr = (d["train_indptr"], d["train_indices"])
    te = (d["test_indptr"], d["test_indices"])
    metric, K = [1, 2, 4], 50
    a = _run_fused(torch, ctx, d["user_emb"], d["item_emb"], None, tr, te, metric, K, "3xtf32")
    b = _run_fused(torch, ctx, d["user_emb"], d["item_emb"], None, tr, te, metric, K, "fp32")
    U = d["users"]
    assert np.max(np.abs(a[3] / U - b[3] / U)) <= TOL_METRIC
    differ = a[0] != b[0]
    assert differ.mean() < 0.01
    if differ.any():
        assert np.max(np.abs(a[1][differ] - b[1][differ])) < TOL_NEAR_TIE
    # oracle on a sample of users
    sample = np.arange(0, U, U // 512)[:512]
    S = oracle.scores(d["user_emb"][sample], d["item_emb"], None)
    sp, si = oracle.dicts_to_csr(sample.tolist(), d["train"])
    oracle.mask_rows(S, sp, si)
    ep, ei = oracle.dicts_to_csr(sample.tolist(), d["test"], dedup_sort=True)
    eper, etop = oracle.eval_scores(S, ep, ei, metric, K, return_topk=True)
    got = a[0][sample]
    diff = got != etop
    if diff.any():
        gs = np.take_along_axis(S, got.astype(np.int64), 1)
        es = np.take_along_axis(S, etop.astype(np.int64), 1)
        assert np.max(np.abs(gs[diff] - es[diff])) < TOL_NEAR_TIE
    assert np.max(np.abs(a[2][sample].astype(np.float64).mean(0) - eper.astype(np.float64).mean(0))) <= TOL_METRIC
    # hits are possible at all (planted test items): the check is not vacuous
    assert eper[:, K - 1].mean() > 0.005


def test_c2_item_permutation_invariance(torch_cuda, ctx, c2_data):
    """Size-independent property: relabelling the items (permuting the item table, the train and
    the test ids consistently) leaves every metric unchanged up to tie order."""
    torch = torch_cuda
    d = c2_data
    U = 4096
    g = np.random.default_rng(0)
    I = d["items"]
    perm = g.permutation(I).astype(np.int32)       # new id of old item j
    inv = np.empty(I, np.int64); inv[perm] = np.arange(I)
    ue = d["user_emb"][:U]
    trp = d["train_indptr"][:U + 1]; tri = d["train_indices"][:trp[-1]]
    tep = d["test_indptr"][:U + 1]; tei = d["test_indices"][:tep[-1]]
    a = _run_fused(torch, ctx, ue, d["item_emb"], None, (trp, tri), (tep, tei), [2, 4], 50, "3xtf32")
    b = _run_fused(torch, ctx, ue, np.ascontiguousarray(d["item_emb"][inv]), None, (trp, perm[tri]), (tep, perm[tei]), [2, 4], 50, "3xtf32")
    assert np.max(np.abs(a[3] - b[3]) / U) <= TOL_METRIC
    same = (perm[a[0]] == b[0])
    assert same.mean() > 0.995


# ---- item-sharded evaluation: per-shard lists + merge == unsharded --------------------------
def _col_partition(tr, lo, hi):
    """Column partition [lo, hi) of a CSR with shard-local ids."""
    indptr, indices = tr
    keep = (indices >= lo) & (indices < hi)
    rows = np.repeat(np.arange(indptr.size - 1), np.diff(indptr))
    cnt = np.bincount(rows[keep], minlength=indptr.size - 1)
    ptr = np.zeros(indptr.size, np.int64)
    np.cumsum(cnt, out=ptr[1:])
    return ptr, (indices[keep] - lo).astype(np.int32)


@pytest.mark.parametrize("precision", ["3xtf32", "fp32"])
@pytest.mark.parametrize("n_shards", [2, 3])
def test_item_sharded_lists_merge_to_the_unsharded_result(torch_cuda, precision, n_shards):
    from skrec_b200 import _native, dist
    torch = torch_cuda
    U, I, d, K = 389, 4000, 64, 50
    ue, ie, b, tr, te = _fused_case(77, U, I, d, True, 80)
    metric = [1, 2, 3, 4, 5]
    ref_ctx = _native.Context(0)
    ref = _run_fused(torch, ref_ctx, ue, ie, b, tr, te, metric, K, precision)
    ued, ied, bd = torch.from_numpy(ue).cuda(), torch.from_numpy(ie).cuda(), torch.from_numpy(b).cuda()
    keys_all = torch.empty((n_shards, U, K), dtype=torch.int64, device="cuda")
    for s in range(n_shards):
        lo, hi = dist.shard_range(I, s, n_shards)
        c = _native.Context(0)
        ptr, idx = _col_partition(tr, lo, hi)
        c.set_train_csr(ptr, idx, hi - lo)
        c.topk_fused(ued, ied[lo:hi], bd[lo:hi], 0, lo, K, keys_all[s], precision=precision)
        torch.cuda.synchronize()
        c.close()
    # every per-shard list holds global ids of its own range, ranked by descending key
    ka = keys_all.cpu().numpy().view(np.uint64)
    for s in range(n_shards):
        lo, hi = dist.shard_range(I, s, n_shards)
        items = (~ka[s]).astype(np.uint32)
        assert items.min() >= lo and items.max() < hi
        assert np.all(ka[s][:, :-1] > ka[s][:, 1:])
    MK = len(metric) * K
    idx = torch.empty((U, K), dtype=torch.int32, device="cuda")
    val = torch.empty((U, K), dtype=torch.float32, device="cuda")
    per = torch.empty((U, MK), dtype=torch.float32, device="cuda")
    sums = torch.zeros(MK, dtype=torch.float64, device="cuda")
    ref_ctx.set_test_csr(te[0], te[1], I)
    # merged in two row slices, like two ranks would
    cut = 200
    ref_ctx.eval_merged_topk(keys_all, 0, cut, 0, metric, K, topk_idx=idx[:cut], topk_val=val[:cut], per_user=per[:cut], sums=sums)
    ref_ctx.eval_merged_topk(keys_all, cut, U - cut, cut, metric, K, topk_idx=idx[cut:], topk_val=val[cut:], per_user=per[cut:], sums=sums)
    torch.cuda.synchronize()
    got = (idx.cpu().numpy(), val.cpu().numpy(), per.cpu().numpy(), sums.cpu().numpy())
    # sharding changes neither a score nor an order: bit-identical to the unsharded run, and right vs the oracle
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    assert np.max(np.abs(got[3] - ref[3])) < 1e-9
    _check_fused(*got, ue, ie, b, tr, te, metric, K)
    ref_ctx.close()


def _items_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as td
    from skrec_b200 import RankingEvaluator, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    data = synth.make(users=1500, items=6000, d=64, nnz_train=60000, nnz_test=12000, seed=11, bias=True)
    model = synth.EmbeddingModel(data["user_emb"], data["item_emb"], data["bias"])
    out = {}
    for shard in ("users", "items"):
        ev = RankingEvaluator(data["train"], data["test"], metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], top_k=[10, 50],
                              device=rank, shard=shard, shard_users=True)
        ev.item_chunk_rows = 512  # 1,500 users: three rounds, the last one short -- the overlapped all-gather path
        rep = ev.evaluate(model)
        dev_vec = ev.evaluate_device(model).cpu().numpy()  # [sums | count], already reduced over the ranks
        assert dev_vec[-1] == 1500 and np.max(np.abs((dev_vec[:-1] / dev_vec[-1]).astype(np.float32).reshape(5, 50)[:, [9, 49]].ravel()
                                                     - np.array(list(rep.values()), np.float32))) == 0.0
        out[shard] = np.array(list(rep.values()), np.float32)
        out[shard + "_path"] = ev.last_stats["path"]
    # host item table: 1/world uploaded per rank + NVLink all-gather (the default above) == every rank uploading all of it;
    # a single evaluated user leaves every rank but one without users -- they still take part in the all-gather
    ev_r = RankingEvaluator(data["train"], data["test"], metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], top_k=[10, 50],
                            device=rank, shard_users=True, upload="replicated")
    out["users_replicated"] = np.array(list(ev_r.evaluate(model).values()), np.float32)
    one = [next(iter(data["test"].keys()))]
    ev_s = RankingEvaluator(data["train"], data["test"], metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], top_k=[10, 50],
                            device=rank, shard_users=True)
    out["one_sharded"] = np.array(list(ev_s.evaluate(model, test_users=one).values()), np.float32)
    out["one_replicated"] = np.array(list(ev_r.evaluate(model, test_users=one).values()), np.float32)
    # one-shot NVLink all-reduce against NCCL: same bits (two ranks: a + b either way), repeated calls (parity, sequence)
    from skrec_b200 import dist
    comm = dist.nvlink_comm(rank, None)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    nv_ok = comm is not None
    worst = 0.0
    if nv_ok:
        for it in range(50):
            v = torch.randn(501, generator=g, device="cuda", dtype=torch.float64)
            a, b = v.clone(), v.clone()
            comm.allreduce(a)
            td.all_reduce(b)
            if world == 2:
                assert torch.equal(a, b), it
            worst = max(worst, float((a - b).abs().max()))
        comm.status()
    if rank == 0:
        np.savez(out_path, users=out["users"], items=out["items"], path=out["items_path"], nvlink=nv_ok, nvlink_err=worst,
                 users_replicated=out["users_replicated"], one_sharded=out["one_sharded"], one_replicated=out["one_replicated"])
    td.destroy_process_group()


def test_item_sharded_evaluator_two_gpus_nccl(torch_cuda, tmp_path):
    """All GPUs of the box (2..8 ranks), NCCL: shard='items' (all-gather + merge) == shard='users' (all-reduce only) == oracle."""
    import socket
    import torch.multiprocessing as mp
    from skrec_b200 import synth
    world = min(torch_cuda.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "rep.npz")
    mp.spawn(_items_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    assert bool(got["nvlink"]) and float(got["nvlink_err"]) < 1e-12  # GPUs of one box: the one-shot path must come up
    assert str(got["path"]).startswith("items:")  # 6,000 / world item rows per shard: below 3,072 the exact FP32 kernel takes them
    assert np.max(np.abs(got["users"] - got["items"])) <= 1e-7
    assert np.array_equal(got["users"], got["users_replicated"])  # same table on the device either way: same bits
    assert np.array_equal(got["one_sharded"], got["one_replicated"]) and np.all(np.isfinite(got["one_sharded"]))
    data = synth.make(users=1500, items=6000, d=64, nnz_train=60000, nnz_test=12000, seed=11, bias=True)
    plain = synth.PredictOnlyModel(data["user_emb"], data["item_emb"], data["bias"])
    per, _ = oracle.evaluate_dicts(plain.predict, data["train"], data["test"], [1, 2, 3, 4, 5], 50)
    expect = oracle.mean_f32(per).reshape(5, 50)[:, np.array([10, 50]) - 1].ravel()
    assert np.max(np.abs(got["items"] - expect)) <= TOL_METRIC


# ---- rows next to the path: pyx_sort-style top-k, adapters, grouped evaluation, CSR ingestion ---
def test_top_k_and_arg_top_k_like_pyx_sort(torch_cuda):
    from skrec_b200 import arg_top_k, top_k
    g = np.random.default_rng(3)
    a = _tie_free(g, 37, 3001)
    exp_idx = np.argsort(-a, axis=1, kind="stable")[:, :20]
    assert np.array_equal(arg_top_k(a, 20, 4), exp_idx)
    assert np.array_equal(top_k(a, 20), np.take_along_axis(a, exp_idx, 1))
    v = a[5]
    assert np.array_equal(arg_top_k(v, 7), exp_idx[5, :7]) and top_k(v, 7).shape == (7,)
    ties = np.array([[1, 5, 5, 2, 5, 0]], np.int32)
    assert arg_top_k(ties, 4).tolist() == [[1, 2, 4, 3]]  # equal values: lower index first
    assert top_k(ties, 4).tolist() == [[5, 5, 5, 2]] and top_k(ties, 4).dtype == np.int32
    ref = oracle.ref_module("pyx_sort")  # the compiled reference (pyx_sort.pyx:151-187) agrees on tie-free input
    if ref is not None:
        assert np.array_equal(ref.pyx_arg_top_k(a, 20, 4), arg_top_k(a, 20, 4))
        assert np.array_equal(ref.pyx_top_k(a, 20, 4), top_k(a, 20, 4))
    with pytest.raises(TypeError):
        arg_top_k(a.astype(np.float64), 3)
    with pytest.raises(ValueError):
        arg_top_k(np.zeros((2, 2, 2), np.float32), 1)


def test_adapters_through_the_evaluator(torch_cuda):
    from skrec_b200 import RankingEvaluator, adapters, synth
    data = synth.make(users=400, items=4000, d=32, nnz_train=9000, nnz_test=2500, seed=21, bias=True)
    metric, top_k = ["Precision", "Recall", "MAP", "NDCG", "MRR"], [5, 20]
    ids = [synth.METRIC_IDS[m] for m in metric]
    cols = np.array(top_k) - 1

    def expect(score_fn):
        per, _ = oracle.evaluate_dicts(score_fn, data["train"], data["test"], ids, max(top_k))
        return oracle.mean_f32(per).reshape(len(ids), max(top_k))[:, cols].ravel()

    ue, ie, b = data["user_emb"], data["item_emb"], data["bias"]
    half = ue.shape[1] // 2
    cases = {
        "dot": (adapters.dot_product(ue, ie, b), lambda us: ue[us] @ ie.T + b),
        "two_tower": (adapters.two_tower_sum(ue[:, :half], ie[:, :half], ue[:, half:], ie[:, half:]), lambda us: ue[us] @ ie.T),
        "cml": (adapters.neg_euclidean(ue, ie), lambda us: -np.linalg.norm(ue[us][:, None, :].astype(np.float64) - ie[None].astype(np.float64), axis=-1).astype(np.float32)),
    }
    for name, (scorer, fn) in cases.items():
        for dev_tables in (False, True):
            if dev_tables:
                scorer.user_table, scorer.item_table = scorer.user_table.cuda(), scorer.item_table.cuda()
                scorer.bias = None if scorer.bias is None else scorer.bias.cuda()
            ev = RankingEvaluator(data["train"], data["test"], metric=metric, top_k=top_k, device=0)
            got = np.array(list(ev.evaluate(scorer).values()), np.float32)
            assert ev.last_stats["path"].startswith("fused:"), name
            assert np.max(np.abs(got - expect(fn))) <= TOL_METRIC, (name, dev_tables)


def test_evaluate_groups_equals_one_evaluation_per_group(torch_cuda):
    from skrec_b200 import RankingEvaluator, adapters, synth
    data = synth.make(users=600, items=3000, d=64, nnz_train=15000, nnz_test=4000, seed=31, bias=False)
    users = list(data["test"].keys())
    deg = np.array([len(data["train"].get(u, ())) for u in users])
    order = np.argsort(deg, kind="stable")
    groups = [[users[i] for i in part] for part in np.array_split(order, 4)] + [[10 ** 9], users[:5] + users[:5]]
    for model in (adapters.dot_product(data["user_emb"], data["item_emb"]), synth.PredictOnlyModel(data["user_emb"], data["item_emb"], None)):
        ev = RankingEvaluator(data["train"], data["test"], metric=["Recall", "NDCG"], top_k=[10, 20], device=0, batch_size=128)
        reports = ev.evaluate_groups(model, groups)
        assert len(reports) == len(groups)
        for g, rep in zip(groups, reports):
            one = ev.evaluate(model, test_users=g)
            a, b_ = np.array(list(rep.values())), np.array(list(one.values()))
            assert rep.metrics_str == one.metrics_str
            if len([u for u in g if u in data["test"]]) == 0:
                continue
            assert np.max(np.abs(a - b_)) <= 2e-7


def test_from_csr_equals_dict_construction(torch_cuda):
    import scipy.sparse as sp
    from skrec_b200 import RankingEvaluator, adapters, synth
    data = synth.make(users=500, items=2000, d=64, nnz_train=12000, nnz_test=3000, seed=41, bias=True)

    def to_csr(d):
        rows = np.concatenate([np.full(len(v), u) for u, v in d.items()])
        cols = np.concatenate(list(d.values()))
        return sp.csr_matrix((np.ones(rows.size, np.float32), (rows, cols)), shape=(data["users"], data["items"]))
    model = adapters.dot_product(data["user_emb"], data["item_emb"], data["bias"])
    kw = dict(metric=["Precision", "NDCG", "MRR"], top_k=[10, 50], device=0)
    a = RankingEvaluator(data["train"], data["test"], **kw).evaluate(model)
    b = RankingEvaluator.from_csr(to_csr(data["train"]), to_csr(data["test"]), **kw).evaluate(model)
    assert np.array_equal(np.array(list(a.values())), np.array(list(b.values())))
    sub = list(data["test"].keys())[5:200:3]
    a = RankingEvaluator(data["train"], data["test"], **kw).evaluate(model, test_users=sub)
    b = RankingEvaluator.from_csr(to_csr(data["train"]), to_csr(data["test"]), **kw).evaluate(model, test_users=sub)
    assert np.array_equal(np.array(list(a.values())), np.array(list(b.values())))


# ---- precision "tf32r": one TF32 pass + exact re-scoring == the FP32 path, bit for bit -----------------------
@pytest.mark.parametrize("prec", ["tf32r", "f16r"])
@pytest.mark.parametrize("U,I,d,bias,K,max_train", FUSED_SHAPES + [(700, 9000, 64, True, 50, 60)])
def test_fused_tf32r_equals_fp32_path_bit_for_bit(torch_cuda, ctx, U, I, d, bias, K, max_train, prec):
    ue, ie, b, tr, te = _fused_case(U + I + 1, U, I, d, bias, max_train)
    metric = [1, 2, 3, 4, 5]
    got = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, prec)
    assert ctx.last_fused_kernel == "tcgen05_" + prec
    ref = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, "fp32")
    assert ctx.last_fused_kernel == "simt_fp32"
    assert np.array_equal(got[0], ref[0])      # same items in the same order
    assert np.array_equal(got[1], ref[1])      # same float32 scores (the exact FMA chain, not the TF32 value)
    assert np.array_equal(got[2], ref[2])      # same per-user metric vectors
    assert np.max(np.abs(got[3] - ref[3])) < 1e-9
    _check_fused(*got, ue, ie, b, tr, te, metric, K)


@pytest.mark.parametrize("prec", ["tf32r", "f16r"])
@pytest.mark.parametrize("d,bias", [(32, True), (64, False), (96, True), (128, True), (128, False), (50, False)])
def test_fused_tf32r_with_working_thresholds_equals_fp32_path(torch_cuda, ctx, d, bias, prec):
    """Catalogue large enough for the sampled thresholds to settle (nearly) every row in the candidate path -- small
    shapes go through the exact fallback and would not exercise the threshold MMA (d <= 96) or the FADD epilogue (d = 128)."""
    U, I, K = 300, 16384, 50
    ue, ie, b, tr, te = _fused_case(1000 + d, U, I, d, bias, 40)
    metric = [1, 2, 4]
    got = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, prec)
    assert ctx.last_fused_kernel == "tcgen05_" + prec
    assert ctx.fused_stats()["exact_rows"] <= U // 20
    if d % 4 != 0:  # the FP32 tile kernel needs d % 4 == 0: the oracle is the judge
        _check_fused(*got, ue, ie, b, tr, te, metric, K, tol_score=TOL_SCORE * max(1.0, 0.5 * np.sqrt(d)))
        return
    ref = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, "fp32")
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    x3 = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, "3xtf32")
    assert ctx.fused_stats()["exact_rows"] <= U // 20
    _check_fused(*x3, ue, ie, b, tr, te, metric, K, tol_score=TOL_SCORE * max(1.0, 0.5 * np.sqrt(d)))


def test_tf32_error_stays_inside_the_band_tf32r_assumes(torch_cuda, ctx):
    """|s_tf32 - s_fp32| of the single-pass kernel (threshold MMA and the epilogue's add-back included) against eps
    of k_sample_thr (k_fused_tc.cuh), on the candidates the kernel reports: the bound must hold with slack."""
    g = np.random.default_rng(5)
    for d, scale in ((32, 0.1), (64, 1.0), (96, 0.5), (128, 3.0)):
        U, I, K = 128, 16384, 50
        ue = (g.standard_normal((U, d)) * scale).astype(np.float32)
        ie = (g.standard_normal((I, d)) * scale).astype(np.float32)
        te = _rand_csr(g, U, I, 5, min_n=1)
        out1 = _run_fused(torch_cuda, ctx, ue, ie, None, None, te, [1], K, "1xtf32")
        assert ctx.fused_stats()["exact_rows"] == 0  # the approximate scores really come from the single pass
        exact = ue.astype(np.float64) @ ie.astype(np.float64).T
        s2 = np.take_along_axis(exact, out1[0].astype(np.int64), 1)
        err = np.abs(out1[1].astype(np.float64) - s2).max(axis=1)
        coef = 2.0 ** -10 + (2.5 * d + 8.0) * 2.0 ** -22
        eps = 1.25 * coef * np.linalg.norm(ue.astype(np.float64), axis=1) * np.linalg.norm(ie.astype(np.float64), axis=1).max()
        assert np.all(err <= 0.5 * eps), (d, float((err / eps).max()))
        assert err.max() > 2.0 ** -20 * np.abs(s2).max()  # the single pass really is inexact: the band is needed


@pytest.mark.parametrize("prec", ["tf32r", "f16r"])
def test_tf32r_ties_and_degenerate_rows_fall_back_exactly(torch_cuda, ctx, prec):
    """Many equal scores (popularity-style integer scores) and rows with fewer unmasked items than K."""
    g = np.random.default_rng(6)
    U, I, d, K = 200, 1500, 64, 20
    ue = np.zeros((U, d), np.float32)
    ue[:, 0] = 1.0
    ie = np.zeros((I, d), np.float32)
    ie[:, 0] = g.integers(0, 12, size=I).astype(np.float32)  # heavy ties
    tr = _rand_csr(g, U, I, 30)
    te = _rand_csr(g, U, I, 10, min_n=1)
    got = _run_fused(torch_cuda, ctx, ue, ie, None, tr, te, [1, 2, 4], K, prec)
    ref = _run_fused(torch_cuda, ctx, ue, ie, None, tr, te, [1, 2, 4], K, "fp32")
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2])
    S = oracle.scores(ue, ie, None)
    oracle.mask_rows(S, tr[0], tr[1])
    eper, etop = oracle.eval_scores(S, te[0], te[1], [1, 2, 4], K, return_topk=True)
    assert np.array_equal(got[0], etop) and np.array_equal(got[2], eper)  # lower item id first among equals


@pytest.mark.parametrize("d,bias_scale", [(64, 0.0), (128, 0.05), (128, 30.0), (64, 1e4)])
@pytest.mark.parametrize("su,si", [(1.0, 1.0), (1e-6, 1e-3), (3e3, 2e2), (1e-12, 1e6)])
def test_f16r_power_of_two_scaling_is_invisible(torch_cuda, ctx, d, bias_scale, su, si):
    """FP16 operands need the tables scaled into fp16's range (per user row, per catalogue): tiny, huge and mixed
    magnitudes, user rows 10^4 apart, a bias that dwarfs the dot products (the threshold operand then forces a smaller
    row scale, or the row is handed to the exact kernel) -- always the FP32 path's items, scores and metrics, bit for bit."""
    g = np.random.default_rng(int(d + 1000 * bias_scale) + 7)
    U, I, K = 260, 16384, 50
    ue = (g.standard_normal((U, d)) * 0.1 * su).astype(np.float32)
    ue[::3] *= 1e-4                                            # rows of very different magnitude: the scale is per row
    ue[1, : d // 2] *= 1e-7                                    # elements far below the row's largest: fp16 subnormals
    ie = (g.standard_normal((I, d)) * 0.1 * si).astype(np.float32)
    ie[::11] *= 30.0
    ie[5, :] *= 1e-9
    b = (g.standard_normal(I) * bias_scale * su * si).astype(np.float32) if bias_scale > 0 else None
    tr = _rand_csr(g, U, I, 40)
    te = _rand_csr(g, U, I, 6, min_n=1)
    metric = [1, 2, 4]
    got = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, "f16r")
    assert ctx.last_fused_kernel == "tcgen05_f16r"
    stats = ctx.fused_stats()
    ref = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, "fp32")
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    if bias_scale <= 30.0:  # a bias 10^4 times the dot products may cost rows to the exact kernel; anything milder must not
        assert stats["exact_rows"] <= U // 10, stats
    if b is not None:  # the bias inside the contraction (default) and added by the epilogue: the same bits
        ctx.set_option("no_aug", 1)
        try:
            alt = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, "f16r")
        finally:
            ctx.set_option("no_aug", 0)
        assert np.array_equal(alt[0], ref[0]) and np.array_equal(alt[1], ref[1]) and np.array_equal(alt[2], ref[2])


def test_c2_full_size_tf32r_equals_fp32_path(torch_cuda, ctx, c2_data):
    d = c2_data
    tr = (d["train_indptr"], d["train_indices"])
    te = (d["test_indptr"], d["test_indices"])
    a = _run_fused(torch_cuda, ctx, d["user_emb"], d["item_emb"], None, tr, te, [1, 2, 4], 50, "tf32r")
    stats = ctx.fused_stats()
    b = _run_fused(torch_cuda, ctx, d["user_emb"], d["item_emb"], None, tr, te, [1, 2, 4], 50, "fp32")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert stats["exact_rows"] < 0.01 * d["users"]  # the fast path settles (almost) every row
    h = _run_fused(torch_cuda, ctx, d["user_emb"], d["item_emb"], None, tr, te, [1, 2, 4], 50, "f16r")  # FP16 operands: the same bits
    assert ctx.last_fused_kernel == "tcgen05_f16r" and ctx.fused_stats()["exact_rows"] < 0.01 * d["users"]
    assert np.array_equal(h[0], b[0]) and np.array_equal(h[1], b[1]) and np.array_equal(h[2], b[2])


def test_c4_catalogue_size_tf32r_equals_fp32_path_and_oracle_sample(torch_cuda, ctx):
    """BASELINE.json configs[3] at the full catalogue size (1M items, d=128, bias, top-100, five metrics) on a 2,048-user
    slice of one GPU's share: the tensor-core path (TF32 candidates + exact re-scoring) gives the FP32 kernel's items,
    scores and per-user metric vectors bit for bit, and both agree with the oracle on a 32-user sample."""
    from skrec_b200 import synth
    torch = torch_cuda
    cfg = dict(synth.CONFIGS["c4"])
    U = 2048
    cfg.update(users=U, nnz_train=U * 50, nnz_test=U * 10)
    d = synth.make(device="cuda", **cfg)
    tr = (d["train_indptr"], d["train_indices"])
    te = (d["test_indptr"], d["test_indices"])
    metric, K = [1, 2, 3, 4, 5], 100
    a = _run_fused(torch, ctx, d["user_emb"], d["item_emb"], d["bias"], tr, te, metric, K, "tf32r")
    assert ctx.last_fused_kernel == "tcgen05_tf32r"
    stats = ctx.fused_stats()
    b = _run_fused(torch, ctx, d["user_emb"], d["item_emb"], d["bias"], tr, te, metric, K, "fp32")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert stats["exact_rows"] < 0.01 * U  # the sampled thresholds settle (almost) every row
    h = _run_fused(torch, ctx, d["user_emb"], d["item_emb"], d["bias"], tr, te, metric, K, "f16r")  # FP16 operands: the same bits
    assert ctx.last_fused_kernel == "tcgen05_f16r" and ctx.fused_stats()["exact_rows"] < 0.01 * U
    assert np.array_equal(h[0], b[0]) and np.array_equal(h[1], b[1]) and np.array_equal(h[2], b[2])
    # every list: K distinct, unmasked items in descending (score, -id) order
    assert np.all((a[1][:, :-1] > a[1][:, 1:]) | ((a[1][:, :-1] == a[1][:, 1:]) & (a[0][:, :-1] < a[0][:, 1:])))
    rows = np.repeat(np.arange(U), np.diff(tr[0]))
    train_keys = rows.astype(np.int64) * d["items"] + tr[1]
    list_keys = (np.arange(U, dtype=np.int64)[:, None] * d["items"] + a[0]).ravel()
    assert not np.isin(list_keys, train_keys).any()
    # oracle on a sample of users
    sample = np.arange(0, U, U // 32)[:32]
    S = oracle.scores(d["user_emb"][sample], d["item_emb"], d["bias"])
    sp, si = oracle.dicts_to_csr(sample.tolist(), d["train"])
    oracle.mask_rows(S, sp, si)
    ep, ei = oracle.dicts_to_csr(sample.tolist(), d["test"], dedup_sort=True)
    eper, etop = oracle.eval_scores(S, ep, ei, metric, K, return_topk=True)
    got = a[0][sample]
    diff = got != etop
    if diff.any():
        gs = np.take_along_axis(S, got.astype(np.int64), 1)
        es = np.take_along_axis(S, etop.astype(np.int64), 1)
        assert np.max(np.abs(gs[diff] - es[diff])) < TOL_NEAR_TIE
    assert diff.mean() < 0.02
    assert np.max(np.abs(a[2][sample].astype(np.float64).mean(0) - eper.astype(np.float64).mean(0))) <= TOL_METRIC
    same = ~diff.any(axis=1)
    assert np.array_equal(a[2][sample][same], eper[same])
    assert eper[:, 4 * K - 1].mean() > 0.005  # NDCG@100 of the planted test items: the check is not vacuous


# ---- randomized shape sweep: every fused precision against the oracle on ragged / awkward shapes -------------
def _sweep_cases():
    g = np.random.default_rng(2026)
    cases = []
    for _ in range(14):
        U = int(g.choice([1, 2, 31, 127, 128, 129, 255, 300, 513]))
        I = int(g.choice([130, 200, 257, 1000, 2049, 4100]))
        d = int(g.choice([1, 3, 4, 17, 32, 50, 64, 100, 128]))
        K = int(g.choice([1, 2, 5, 10, 50, 64, 65, 100, 128]))
        cases.append((U, I, d, min(K, I // 2), bool(g.integers(2)), int(g.choice([0, 1, 5, 40])), int(g.integers(1 << 30))))
    return cases


@pytest.mark.parametrize("U,I,d,K,bias,max_train,seed", _sweep_cases())
def test_fused_precisions_on_awkward_shapes(torch_cuda, ctx, U, I, d, K, bias, max_train, seed):
    g = np.random.default_rng(seed)
    ue = (g.standard_normal((U, d)) * 0.3).astype(np.float32)
    ie = (g.standard_normal((I, d)) * 0.3).astype(np.float32)
    b = (g.standard_normal(I) * 0.05).astype(np.float32) if bias else None
    tr = _rand_csr(g, U, I, max_train) if max_train > 0 else None
    te = _rand_csr(g, U, I, 8, min_n=0)  # users without test items: all metrics 0, still counted by the caller
    metric = [int(x) for x in g.permutation(5)[: int(g.integers(1, 6))] + 1]
    precisions = ["3xtf32", "tf32r", "f16r"] + (["fp32"] if d % 4 == 0 else [])
    outs = {}
    for prec in precisions:
        outs[prec] = _run_fused(torch_cuda, ctx, ue, ie, b, tr, te, metric, K, prec)
        # scores here reach |s| ~ 0.09 sqrt(d) * 4: the score tolerance scales with their magnitude
        _check_fused(*outs[prec], ue, ie, b, tr, te, metric, K, tol_score=TOL_SCORE * max(1.0, 0.5 * d ** 0.5))
    if "fp32" in outs:
        assert np.array_equal(outs["tf32r"][0], outs["fp32"][0]) and np.array_equal(outs["tf32r"][2], outs["fp32"][2])
        assert np.array_equal(outs["f16r"][0], outs["fp32"][0]) and np.array_equal(outs["f16r"][2], outs["fp32"][2])


def test_fused_special_values_and_limits(torch_cuda, ctx):
    """Zero vectors (every score ties at 0), huge and tiny magnitudes, -0.0; errors for shapes the path cannot take."""
    from skrec_b200 import _native
    g = np.random.default_rng(9)
    U, I, d, K = 140, 900, 64, 10
    ue = (g.standard_normal((U, d))).astype(np.float32)
    ie = (g.standard_normal((I, d))).astype(np.float32)
    ue[:5] = 0.0                      # all-tie rows: ids 0..K-1 in order (minus train items)
    ue[5:10] *= 1e18                  # large scores (|s| ~ 1e19), still finite
    ue[10:15] *= 1e-18                # tiny scores
    ie[::7] = -0.0
    tr = _rand_csr(g, U, I, 20)
    te = _rand_csr(g, U, I, 6, min_n=1)
    for prec in ("3xtf32", "tf32r", "f16r", "fp32"):
        idx, val, per, sums = _run_fused(torch_cuda, ctx, ue, ie, None, tr, te, [1, 2, 3, 4, 5], K, prec)
        S = oracle.scores(ue, ie, None)
        oracle.mask_rows(S, tr[0], tr[1])
        eper, etop = oracle.eval_scores(S, te[0], te[1], [1, 2, 3, 4, 5], K, return_topk=True)
        assert np.array_equal(idx[:5], etop[:5]), prec       # exact ties: lower item id first
        assert np.array_equal(per[:5], eper[:5]), prec
        rel = np.abs(np.take_along_axis(S, idx.astype(np.int64), 1) - np.take_along_axis(S, etop.astype(np.int64), 1))
        scale = np.abs(np.take_along_axis(S, etop.astype(np.int64), 1)) + 1e-30
        assert np.max(rel / scale) < 1e-5, prec               # any rank difference is a near-tie relative to the score scale
        assert np.all(np.isfinite(sums))
    u, i = torch_cuda.zeros((4, 200), device="cuda"), torch_cuda.zeros((50, 200), device="cuda")
    ctx.set_train_csr(None, None, 50)
    ctx.set_test_csr(np.array([0, 1, 2, 3, 4], np.int64), np.array([0, 1, 2, 3], np.int32), 50)
    with pytest.raises(_native.NativeError):   # d > 128 has no tensor-core instantiation: explicit, not a silent fallback
        ctx.eval_fused(u, i, None, 0, [1], 5, precision="3xtf32", sums=torch_cuda.zeros(5, dtype=torch_cuda.float64, device="cuda"))
    ctx.eval_fused(u, i, None, 0, [1], 5, precision="auto", sums=torch_cuda.zeros(5, dtype=torch_cuda.float64, device="cuda"))
    assert ctx.last_fused_kernel == "simt_fp32"
    with pytest.raises(_native.NativeError):   # n_items < top_k (evaluate.h:45 would read out of bounds)
        ctx.eval_fused(u, i, None, 0, [1], 51, sums=torch_cuda.zeros(51, dtype=torch_cuda.float64, device="cuda"))


def test_evaluator_routes_shapes_outside_the_fused_kernels(torch_cuda):
    """top-K > 128 or a wide d that is not a multiple of 4: the library's own FP32 tile kernels (score blocks + the
    score-matrix kernels for top-K > 128; padded rows + the fused FP32 kernel for the odd width) -- no library GEMM."""
    from skrec_b200 import RankingEvaluator, adapters, synth
    import torch
    for d, top_k, path, on_dev in ((64, [10, 200], "fused:simt_fp32_blocks", False), (130, [5, 20], "fused:simt_fp32", True),
                                   (50, [3, 300], "fused:simt_fp32_blocks", True), (64, [10, 200], "fused:simt_fp32_blocks", True)):
        data = synth.make(users=300, items=1500, d=d, nnz_train=6000, nnz_test=1500, seed=51 + d, bias=True)
        metric = ["Precision", "Recall", "MAP", "NDCG", "MRR"]
        ids = [synth.METRIC_IDS[m] for m in metric]
        if on_dev:
            model = adapters.dot_product(torch.from_numpy(data["user_emb"]).cuda(), torch.from_numpy(data["item_emb"]).cuda(),
                                         torch.from_numpy(data["bias"]).cuda())
        else:
            model = adapters.dot_product(data["user_emb"], data["item_emb"], data["bias"])
        ev = RankingEvaluator(data["train"], data["test"], metric=metric, top_k=top_k, device=0, batch_size=128)
        got = np.array(list(ev.evaluate(model).values()), np.float32)
        assert ev.last_stats["path"] == path, ev.last_stats["path"]
        per, _ = oracle.evaluate_dicts(synth.PredictOnlyModel(data["user_emb"], data["item_emb"], data["bias"]).predict,
                                       data["train"], data["test"], ids, max(top_k))
        expect = oracle.mean_f32(per).reshape(len(ids), max(top_k))[:, np.array(top_k) - 1].ravel()
        assert np.max(np.abs(got - expect)) <= TOL_METRIC


def test_evaluator_routes_small_catalogues_by_requested_rank(torch_cuda):
    """ml-1m-sized catalogue: top-20 fits the sampled-threshold plan (tensor cores); top-50 and the reference's default
    top-100 (run_config.py:16) do not: the exact FP32 fused kernel takes them -- a `fused:` path, no library GEMM."""
    from skrec_b200 import RankingEvaluator, adapters, synth
    data = synth.make(users=500, items=3706, d=64, nnz_train=20000, nnz_test=2500, seed=77, bias=True)
    metric = ["Precision", "Recall", "NDCG"]
    ids = [synth.METRIC_IDS[m] for m in metric]
    model = adapters.dot_product(data["user_emb"], data["item_emb"], data["bias"])
    for top_k, path in (([10, 20], "fused:tcgen05"), ([20, 50], "fused:simt_fp32"), (list(range(10, 101, 10)), "fused:simt_fp32")):
        ev = RankingEvaluator(data["train"], data["test"], metric=metric, top_k=top_k, device=0)
        got = np.array(list(ev.evaluate(model).values()), np.float32)
        assert ev.last_stats["path"].startswith(path), ev.last_stats["path"]
        per, _ = oracle.evaluate_dicts(synth.PredictOnlyModel(data["user_emb"], data["item_emb"], data["bias"]).predict,
                                       data["train"], data["test"], ids, max(top_k))
        expect = oracle.mean_f32(per).reshape(len(ids), max(top_k))[:, np.array(top_k) - 1].ravel()
        assert np.max(np.abs(got - expect)) <= TOL_METRIC


@pytest.mark.parametrize("U,I,seg_rows", [(700, 20000, -1), (700, 20000, 0), (3000, 5000, 100)])
def test_exact_fallback_when_thresholds_are_useless(torch_cuda, U, I, seg_rows):
    """rank = 1 makes the sampled threshold the largest sampled score: most rows collect fewer than K candidates and
    go through the exact fallback -- segmented (few rows), whole-row (many rows) or both.  Results stay exact.
    (This shape also reproduced a barrier race in the streaming selection that the default plan never triggered.)"""
    from skrec_b200 import _native, synth
    torch = torch_cuda
    d = synth.make(users=U, items=I, d=64, nnz_train=U * 20, nnz_test=U * 5, seed=3, device="cuda")
    c = _native.Context(0)
    c.set_train_csr(d["train_indptr"], d["train_indices"], I)
    c.set_test_csr(d["test_indptr"], d["test_indices"], I)
    c.set_option("rank", 1)
    c.set_option("sample_tiles", 2)
    c.set_option("exact_seg_rows", seg_rows)
    ue, ie = torch.from_numpy(d["user_emb"]).cuda(), torch.from_numpy(d["item_emb"]).cuda()
    S = oracle.scores(d["user_emb"], d["item_emb"], None)
    oracle.mask_rows(S, d["train_indptr"], d["train_indices"])
    metric, K = [1, 2, 3, 4, 5], 20
    eper, etop = oracle.eval_scores(S, d["test_indptr"], d["test_indices"], metric, K, return_topk=True)
    for prec in ("3xtf32", "tf32r"):
        for rep in range(2):
            idx = torch.empty((U, K), dtype=torch.int32, device="cuda")
            val = torch.empty((U, K), dtype=torch.float32, device="cuda")
            per = torch.empty((U, len(metric) * K), dtype=torch.float32, device="cuda")
            sums = torch.zeros(len(metric) * K, dtype=torch.float64, device="cuda")
            c.eval_fused(ue, ie, None, 0, metric, K, precision=prec, topk_idx=idx, topk_val=val, per_user=per, sums=sums)
            torch.cuda.synchronize()
            assert c.fused_stats()["exact_rows"] > 0.2 * U
            _check_fused(idx.cpu().numpy(), val.cpu().numpy(), per.cpu().numpy(), sums.cpu().numpy(),
                         d["user_emb"], d["item_emb"], None, (d["train_indptr"], d["train_indices"]),
                         (d["test_indptr"], d["test_indices"]), metric, K)
    c.close()


def test_eval_score_matrix_accepts_sets_like_bert4rec(torch_cuda):
    """bert4rec_utils.py:25,78: test_items is a list of Python sets; same numbers as arrays."""
    from skrec_b200 import eval_score_matrix
    g = np.random.default_rng(12)
    s = g.standard_normal((40, 300)).astype(np.float32)
    items = [g.choice(300, size=int(g.integers(1, 6)), replace=False).astype(np.int32) for _ in range(40)]
    a = eval_score_matrix(s.copy(), items, [1, 2, 3, 4, 5], 10, 4)
    b = eval_score_matrix(s.copy(), [set(int(x) for x in it) for it in items], [1, 2, 3, 4, 5], 10, 4)
    assert np.array_equal(a, b)


def test_plan_cache_never_mixes_up_two_subsets_of_the_same_length(torch_cuda):
    """a cache hit is trusted only after comparing the users."""
    from skrec_b200 import RankingEvaluator, synth
    data = synth.make(users=400, items=2000, d=32, nnz_train=8000, nnz_test=2000, seed=5, bias=False)
    model = synth.EmbeddingModel(data["user_emb"], data["item_emb"], None)
    ev = RankingEvaluator(data["train"], data["test"], metric=["Recall", "NDCG"], top_k=[5, 10], device=0, precision="fp32")
    a, b = list(range(0, 200)), list(range(200, 400))
    ra = np.array(list(ev.evaluate(model, a).values()))
    rb = np.array(list(ev.evaluate(model, b).values()))
    # force a key collision: whatever the hash, the second lookup must notice the users differ
    ka = ev._resolve_users(a)[1] + (0, 1, "users")
    plan_b = ev._plans[ev._resolve_users(b)[1] + (0, 1, "users")]
    ev._plans[ka] = plan_b
    ra2 = np.array(list(ev.evaluate(model, a).values()))
    assert np.array_equal(ra, ra2) and not np.array_equal(ra, rb)
    ev2 = RankingEvaluator(data["train"], data["test"], metric=["Recall", "NDCG"], top_k=[5, 10], device=0, precision="fp32")
    assert np.array_equal(np.array(list(ev2.evaluate(model, b).values())), rb)
    for i in range(6):  # the cache is bounded and evicted contexts are closed
        ev.evaluate(model, list(range(i, i + 50)))
    assert len(ev._plans) <= ev._MAX_PLANS


def test_row_chunked_fused_equals_one_call(torch_cuda, ctx):
    """skr_eval_fused cuts large row counts into chunks (item-side preparation once): same per-user block, same sums."""
    from skrec_b200 import _native, synth
    torch = torch_cuda
    d = synth.make(users=1000, items=9000, d=64, nnz_train=30000, nnz_test=6000, seed=21, bias=True, device="cuda")
    ue, ie, b = (torch.from_numpy(d[k]).cuda() for k in ("user_emb", "item_emb", "bias"))
    metric, K = [1, 2, 3, 4, 5], 20
    out = {}
    for rows in (0, 256):
        c = _native.Context(0)
        c.set_train_csr(d["train_indptr"], d["train_indices"], 9000)
        c.set_test_csr(d["test_indptr"], d["test_indices"], 9000)
        c.set_option("chunk_rows", rows)
        for prec in ("tf32r", "3xtf32", "fp32"):
            idx = torch.empty((1000, K), dtype=torch.int32, device="cuda")
            per = torch.empty((1000, len(metric) * K), dtype=torch.float32, device="cuda")
            sums = torch.zeros(len(metric) * K, dtype=torch.float64, device="cuda")
            c.eval_fused(ue, ie, b, 0, metric, K, precision=prec, topk_idx=idx, per_user=per, sums=sums)
            torch.cuda.synchronize()
            out[(rows, prec)] = (idx.cpu().numpy(), per.cpu().numpy(), sums.cpu().numpy())
        c.close()
    for prec in ("tf32r", "3xtf32", "fp32"):
        a, bb = out[(0, prec)], out[(256, prec)]
        assert np.array_equal(a[0], bb[0]) and np.array_equal(a[1], bb[1])
        assert np.max(np.abs(a[2] - bb[2])) < 1e-9


@pytest.mark.parametrize("U,I,d,bias,K", [(1500, 20000, 64, True, 50), (900, 30000, 128, False, 100), (700, 9000, 32, True, 20)])
def test_tf32r_on_heavy_tailed_tables_equals_fp32_and_retries_instead_of_walking_rows(torch_cuda, U, I, d, bias, K):
    """trained tables have heavy-tailed item norms and a few outliers; the single-pass error band
    (proportional to max ||item||) then leaves rows unsettled.  They are retried in three TF32 passes (band 10-20x
    narrower) instead of being walked one by one; results stay bit-identical to the exact FP32 path."""
    from skrec_b200 import _native, synth
    torch = torch_cuda
    dta = synth.make(users=U, items=I, d=d, nnz_train=U * 30, nnz_test=U * 6, seed=100 + d, bias=bias, norms="heavy", device="cuda")
    norms = np.linalg.norm(dta["item_emb"], axis=1)
    assert norms.max() > 6 * np.median(norms)
    c = _native.Context(0)
    c.set_train_csr(dta["train_indptr"], dta["train_indices"], I)
    c.set_test_csr(dta["test_indptr"], dta["test_indices"], I)
    ue, ie = torch.from_numpy(dta["user_emb"]).cuda(), torch.from_numpy(dta["item_emb"]).cuda()
    b = None if dta["bias"] is None else torch.from_numpy(dta["bias"]).cuda()
    metric = [1, 2, 3, 4, 5]
    out = {}
    c.set_option("retry_min", 1)  # always retry (the cost model would hand a handful of rows straight to the exact kernel)
    for prec in ("tf32r", "fp32", "f16r", "tf32r_model"):  # f16r: its unsettled rows take the same three-pass TF32 retry
        if prec == "tf32r_model":
            c.set_option("retry_min", -1)
            prec = "tf32r"
            key = "tf32r_model"
        else:
            key = prec
        idx = torch.empty((U, K), dtype=torch.int32, device="cuda")
        val = torch.empty((U, K), dtype=torch.float32, device="cuda")
        per = torch.empty((U, len(metric) * K), dtype=torch.float32, device="cuda")
        sums = torch.zeros(len(metric) * K, dtype=torch.float64, device="cuda")
        c.eval_fused(ue, ie, b, 0, metric, K, precision=prec, topk_idx=idx, topk_val=val, per_user=per, sums=sums)
        torch.cuda.synchronize()
        out[key] = (idx.cpu().numpy(), val.cpu().numpy(), per.cpu().numpy(), sums.cpu().numpy())
        if key == "tf32r":
            st = c.fused_stats()
    a, f, am = out["tf32r"], out["fp32"], out["tf32r_model"]
    assert np.array_equal(a[0], f[0]) and np.array_equal(a[1], f[1]) and np.array_equal(a[2], f[2])
    assert np.array_equal(am[0], f[0]) and np.array_equal(am[1], f[1]) and np.array_equal(am[2], f[2])
    h = out["f16r"]
    assert np.array_equal(h[0], f[0]) and np.array_equal(h[1], f[1]) and np.array_equal(h[2], f[2])
    assert np.max(np.abs(a[3] - f[3])) < 1e-9 and np.max(np.abs(am[3] - f[3])) < 1e-9
    _check_fused(*a, dta["user_emb"], dta["item_emb"], dta["bias"], (dta["train_indptr"], dta["train_indices"]),
                 (dta["test_indptr"], dta["test_indices"]), metric, K)
    # the retry settles (nearly) everything the first attempt could not: the per-row exact kernel is the exception
    assert st["exact_rows"] <= max(8, st["retried_rows"] // 4), st
    c.close()


def test_three_pass_error_stays_inside_the_band_the_retry_assumes(torch_cuda, ctx):
    """eps3 = 1.25 [(3.25 d + 11) 2^-22 ||u|| max||i|| + 2^-22 max|b|] must bound |s_3xtf32 - s_fp32| with room to spare."""
    g = np.random.default_rng(8)
    for d in (32, 64, 128):
        U, I, K = 256, 4096, 64
        ue = (g.standard_normal((U, d)) * g.lognormal(0, 1, (U, 1))).astype(np.float32)
        ie = (g.standard_normal((I, d)) * g.lognormal(0, 1, (I, 1))).astype(np.float32)
        te = _rand_csr(g, U, I, 3, min_n=1)
        idx3, val3, _, _ = _run_fused(torch_cuda, ctx, ue, ie, None, None, te, [1], K, "3xtf32")
        exact = np.take_along_axis(oracle.scores(ue, ie, None), idx3.astype(np.int64), 1)
        eps3 = 1.25 * (3.25 * d + 11) * 2.0 ** -22 * np.linalg.norm(ue, axis=1, keepdims=True) * np.linalg.norm(ie, axis=1).max()
        assert np.all(np.abs(val3 - exact) <= 0.5 * eps3), d


def test_translation_scorers_neg_l2_plus_bias(torch_cuda):
    """f1 remainder (TransRec.py:86-93, SGAT.py:300): score = -||q - i|| + b_i through the fused FP32 tile kernel with a
    distance inner loop, against the reference expression evaluated in float64 and ranked by the oracle."""
    import torch
    from skrec_b200 import RankingEvaluator, adapters
    g = np.random.default_rng(41)
    for U, I, d, top_k in ((600, 5000, 32, [5, 20]), (300, 2000, 50, [10, 50]), (257, 3000, 64, [10, 200])):
        ue = (g.standard_normal((U, d)) * 0.3).astype(np.float32)
        ie = (g.standard_normal((I, d)) * 0.3).astype(np.float32)
        gt = (g.standard_normal(d) * 0.1).astype(np.float32)
        b = (g.standard_normal(I) * 0.2).astype(np.float32)
        tr = _rand_csr(g, U, I, 25, min_n=1)
        te = _rand_csr(g, U, I, 8, min_n=1)
        train = {u: tr[1][tr[0][u]:tr[0][u + 1]] for u in range(U)}
        test = {u: te[1][te[0][u]:te[0][u + 1]] for u in range(U)}
        last = np.array([train[u][-1] for u in range(U)], np.int64)
        metric = ["Precision", "Recall", "MAP", "NDCG", "MRR"]
        model = adapters.transrec(ue, gt, ie, b, last)
        assert model.score_fn == "neg_l2"
        # the reference expression (TransRec.py:89-92), float64
        q = ue.astype(np.float64) + gt.astype(np.float64) + ie[last].astype(np.float64)
        S = -np.sqrt(((q[:, None, :] - ie[None, :, :].astype(np.float64)) ** 2).sum(-1)) + b.astype(np.float64)
        S = S.astype(np.float32)
        assert np.max(np.abs(model.predict(list(range(U))) - S)) < 1e-5   # the scorer's own predict is that expression
        oracle.mask_rows(S, tr[0], tr[1])
        K = max(top_k)
        per = oracle.eval_scores(S, te[0], te[1], [1, 2, 3, 4, 5], K)
        expect = oracle.mean_f32(per).reshape(5, K)[:, np.array(top_k) - 1].ravel()
        for on_dev in (False, True):
            m = model
            if on_dev:
                m = adapters.neg_l2_plus_bias(torch.from_numpy((ue + gt + ie[last]).astype(np.float32)).cuda(), torch.from_numpy(ie).cuda(),
                                              torch.from_numpy(b).cuda())
            ev = RankingEvaluator(train, test, metric=metric, top_k=top_k, device=0)
            got = np.array(list(ev.evaluate(m).values()), np.float32)
            assert "negl2" in ev.last_stats["path"], ev.last_stats["path"]
            assert np.max(np.abs(got - expect)) <= TOL_METRIC, (U, I, d, on_dev, float(np.max(np.abs(got - expect))))
        # a dot-product model afterwards on the same evaluator: the option does not stick
        ev2 = RankingEvaluator(train, test, metric=metric, top_k=[5], device=0, precision="fp32")
        ev2.evaluate(model)
        ev2.evaluate(adapters.dot_product(ue, ie, b))
        assert ev2.last_stats["path"] == "fused:simt_fp32"


def test_negative_sampler_like_randint_h(torch_cuda):
    """f-5 (randint.h:22-128, random.py:9-41): range, exclusion, distinctness without replacement, uniformity,
    probabilities, reproducibility from the seed, the reference's argument errors and return types."""
    from skrec_b200 import batch_randint_choice, randint_choice
    g = np.random.default_rng(0)
    high = 5000
    excl = [np.unique(g.integers(0, high, size=int(n))) for n in g.integers(1, 400, size=300)]
    size = g.integers(1, 200, size=300)
    a = batch_randint_choice(high, size, replace=True, exclusion=excl, thread_num=4, seed=7)
    b = batch_randint_choice(high, size, replace=True, exclusion=excl, seed=7)
    c = batch_randint_choice(high, size, replace=True, exclusion=excl, seed=8)
    assert isinstance(a, list) and len(a) == 300
    assert all(x.dtype == np.int32 and x.shape == (s,) for x, s in zip(a, size))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and any(not np.array_equal(x, y) for x, y in zip(a, c))
    for x, e in zip(a, excl):
        assert x.min() >= 0 and x.max() < high and not np.isin(x, e).any()
    d = batch_randint_choice(high, size, replace=False, exclusion=excl, seed=9)
    for x, e, s in zip(d, excl, size):
        assert x.shape == (s,) and np.unique(x).size == s and not np.isin(x, e).any() and x.min() >= 0 and x.max() < high
    # nearly exhaustive draw without replacement: 90 of the 100 values left after excluding 28
    ex = np.arange(0, 128, 1)[:28]
    x = randint_choice(128, size=90, replace=False, exclusion=ex, seed=3)
    assert np.unique(x).size == 90 and not np.isin(x, ex).any()
    # uniformity: chi-square of 2*10^6 draws over 1,000 values minus 100 excluded (900 cells, expected 2,222 each)
    ex = np.arange(0, 1000, 10)
    x = randint_choice(1000, size=2_000_000, exclusion=ex, seed=11)
    cnt = np.bincount(x, minlength=1000)
    assert cnt[ex].sum() == 0
    cells = np.delete(cnt, ex)
    chi2 = float(((cells - cells.mean()) ** 2 / cells.mean()).sum())
    assert 700 < chi2 < 1100, chi2   # 899 degrees of freedom: mean 899, sd 42
    # probabilities (one row per element; zero-probability values never come out)
    p = np.zeros((2, 10), np.float32)
    p[0, [1, 3]] = [1.0, 3.0]
    p[1, :] = 1.0
    y = batch_randint_choice(10, [40000, 10], p=p, seed=5)
    assert set(np.unique(y[0])) == {1, 3} and abs((y[0] == 3).mean() - 0.75) < 0.01
    # scalar form: int for size 1, array otherwise; the reference's errors
    assert isinstance(randint_choice(50, seed=1), (int, np.integer))
    assert randint_choice(50, size=7, seed=1).shape == (7,)
    for bad in (lambda: randint_choice(1), lambda: randint_choice(10, size=0), lambda: randint_choice(10, replace=1),
                lambda: randint_choice(10, size=10, replace=False), lambda: randint_choice(10, p=[0.5, 0.5]),
                lambda: batch_randint_choice(10, [[1, 2]]), lambda: batch_randint_choice(10, [3, 0]),
                lambda: batch_randint_choice(10, [3, 3], exclusion=[[1]]),
                lambda: batch_randint_choice(10, [9], replace=False, exclusion=[[1, 2]])):
        with pytest.raises((ValueError, TypeError)):
            bad()
    # the training-side use (data_iterator.py:81-94): negatives of every user against its positives, on the device
    out, ptr = batch_randint_choice(high, size, exclusion=excl, seed=12, as_tensor=True)
    assert out.is_cuda and out.dtype == torch_cuda.int32 and int(ptr[-1]) == int(size.sum())


# ---- the remaining BASELINE.json configs at full size ---------------------------------------------
def _oracle_sample_check(d, bias, a, sample, metric, K, min_hit):
    """oracle on `sample` users of workload d against the fused outputs a = (idx, val, per, sums)"""
    S = oracle.scores(d["user_emb"][sample], d["item_emb"], bias)
    sp, si = oracle.dicts_to_csr(sample.tolist(), d["train"])
    oracle.mask_rows(S, sp, si)
    ep, ei = oracle.dicts_to_csr(sample.tolist(), d["test"], dedup_sort=True)
    eper, etop = oracle.eval_scores(S, ep, ei, metric, K, return_topk=True)
    got = a[0][sample]
    diff = got != etop
    if diff.any():
        gs = np.take_along_axis(S, got.astype(np.int64), 1)
        es = np.take_along_axis(S, etop.astype(np.int64), 1)
        assert np.max(np.abs(gs[diff] - es[diff])) < TOL_NEAR_TIE
    assert diff.mean() < 0.02
    assert np.max(np.abs(a[2][sample].astype(np.float64).mean(0) - eper.astype(np.float64).mean(0))) <= TOL_METRIC
    same = ~diff.any(axis=1)
    assert np.array_equal(a[2][sample][same], eper[same])  # per-user metric vectors bit for bit where the lists agree
    assert eper[:, K - 1].mean() > min_hit                 # planted test items: the check is not vacuous


def test_c1_full_size_every_user_against_the_oracle(torch_cuda, ctx):
    """BASELINE.json configs[0] (ml-1m shape, bias, top-20) whole: every one of the 6,040 users against the oracle, through
    the tensor-core path the evaluator picks at this size (3xTF32) and the exact FP32 path; dense train rows (132 per user)
    exercise the batched mask builder."""
    from skrec_b200 import synth
    d = synth.make_config("c1", device="cuda")
    tr, te = (d["train_indptr"], d["train_indices"]), (d["test_indptr"], d["test_indices"])
    metric, K = [1, 2, 4], 20
    a = _run_fused(torch_cuda, ctx, d["user_emb"], d["item_emb"], d["bias"], tr, te, metric, K, "auto")
    assert ctx.last_fused_kernel == "tcgen05_3xtf32"
    b = _run_fused(torch_cuda, ctx, d["user_emb"], d["item_emb"], d["bias"], tr, te, metric, K, "fp32")
    assert np.max(np.abs(a[3] - b[3])) / d["users"] <= TOL_METRIC
    _oracle_sample_check(d, d["bias"], b, np.arange(d["users"]), metric, K, 0.005)
    _oracle_sample_check(d, d["bias"], a, np.arange(d["users"]), metric, K, 0.005)


@pytest.mark.parametrize("name", ["c3a", "c3b"])
def test_c3_full_size_tf32r_vs_fp32_slice_and_oracle_sample(torch_cuda, ctx, name):
    """BASELINE.json configs[2] at full size: the default path (TF32 candidates + exact re-scoring) on every user; the
    exact FP32 kernel on a 4,096-user slice gives the same items, scores and metric vectors bit for bit; the oracle on a
    256-user sample; every list ordered and free of train items."""
    from skrec_b200 import synth
    d = synth.make_config(name, device="cuda")
    U, I = d["users"], d["items"]
    tr, te = (d["train_indptr"], d["train_indices"]), (d["test_indptr"], d["test_indices"])
    metric, K = [1, 2, 4], 50
    a = _run_fused(torch_cuda, ctx, d["user_emb"], d["item_emb"], None, tr, te, metric, K, "auto")
    assert ctx.last_fused_kernel == "tcgen05_tf32r" and ctx.fused_stats()["exact_rows"] < 0.01 * U
    n = 4096
    trs, tes = (tr[0][:n + 1], tr[1][:tr[0][n]]), (te[0][:n + 1], te[1][:te[0][n]])
    b = _run_fused(torch_cuda, ctx, d["user_emb"][:n], d["item_emb"], None, trs, tes, metric, K, "fp32")
    assert np.array_equal(a[0][:n], b[0]) and np.array_equal(a[1][:n], b[1]) and np.array_equal(a[2][:n], b[2])
    assert np.all((a[1][:, :-1] > a[1][:, 1:]) | ((a[1][:, :-1] == a[1][:, 1:]) & (a[0][:, :-1] < a[0][:, 1:])))
    rows = np.repeat(np.arange(U), np.diff(tr[0]))
    assert not np.isin((np.arange(U, dtype=np.int64)[:, None] * I + a[0]).ravel(), rows.astype(np.int64) * I + tr[1]).any()
    _oracle_sample_check(d, None, a, np.arange(0, U, U // 256)[:256], metric, K, 0.005)


def test_c5_catalogue_size_item_shards_merge_to_the_unsharded_lists(torch_cuda):
    """BASELINE.json configs[4] at the full catalogue size (10^7 items, d = 128, top-100) on 1,024 users: the catalogue cut in
    8 item shards of 1.25M rows (what each of 8 ranks holds), per-shard top-100 rank keys, merge -- bit-identical to the
    unsharded evaluation; the oracle's exact scores on 8 users."""
    from skrec_b200 import _native, dist
    torch = torch_cuda
    U, I, d, K, W = 1024, 10_000_000, 128, 100, 8
    g = torch.Generator(device="cuda").manual_seed(2026)
    ue = torch.randn((U, d), generator=g, device="cuda") * 0.1
    ie = torch.randn((I, d), generator=g, device="cuda") * 0.1
    rng = np.random.default_rng(5)
    tr = (np.arange(U + 1, dtype=np.int64) * 50, rng.integers(0, I, size=U * 50, dtype=np.int32))
    te = (np.arange(U + 1, dtype=np.int64) * 10, rng.integers(0, I, size=U * 10, dtype=np.int32))
    metric = [1, 2, 3, 4, 5]
    ref_ctx = _native.Context(0)
    ref_ctx.set_train_csr(tr[0], tr[1], I)
    ref_ctx.set_test_csr(te[0], te[1], I)
    idx = torch.empty((U, K), dtype=torch.int32, device="cuda")
    val = torch.empty((U, K), dtype=torch.float32, device="cuda")
    per = torch.empty((U, 5 * K), dtype=torch.float32, device="cuda")
    sums = torch.zeros(5 * K, dtype=torch.float64, device="cuda")
    ref_ctx.eval_fused(ue, ie, None, 0, metric, K, topk_idx=idx, topk_val=val, per_user=per, sums=sums)
    torch.cuda.synchronize()
    keys = torch.empty((W, U, K), dtype=torch.int64, device="cuda")
    sh = _native.Context(0)
    for r in range(W):
        lo, hi = dist.shard_range(I, r, W)
        ptr, ind = _col_partition(tr, lo, hi)
        sh.set_train_csr(ptr, ind, hi - lo)
        sh.topk_fused(ue, ie[lo:hi], None, 0, lo, K, keys[r])
    idx2 = torch.empty_like(idx); val2 = torch.empty_like(val); per2 = torch.empty_like(per)
    sums2 = torch.zeros_like(sums)
    ref_ctx.eval_merged_topk(keys, 0, U, 0, metric, K, topk_idx=idx2, topk_val=val2, per_user=per2, sums=sums2)
    torch.cuda.synchronize()
    assert torch.equal(idx, idx2) and torch.equal(val, val2) and torch.equal(per, per2)
    assert float((sums - sums2).abs().max()) < 1e-9
    # exact scores of 8 users (float64 on the device, then the oracle's selection on the host)
    sample = np.arange(0, U, U // 8)[:8]
    S = (ue[sample].double() @ ie.double().T).float().cpu().numpy()
    sp = np.arange(9, dtype=np.int64) * 50
    si = np.concatenate([tr[1][tr[0][u]:tr[0][u + 1]] for u in sample])
    oracle.mask_rows(S, sp, si)
    ep = np.arange(9, dtype=np.int64) * 10
    ei = np.concatenate([np.unique(te[1][te[0][u]:te[0][u + 1]]) for u in sample])
    ep = np.concatenate([[0], np.cumsum([np.unique(te[1][te[0][u]:te[0][u + 1]]).size for u in sample])]).astype(np.int64)
    eper, etop = oracle.eval_scores(S, ep, ei, metric, K, return_topk=True)
    got = idx.cpu().numpy()[sample]
    diff = got != etop
    if diff.any():
        gs = np.take_along_axis(S, got.astype(np.int64), 1)
        es = np.take_along_axis(S, etop.astype(np.int64), 1)
        assert np.max(np.abs(gs[diff] - es[diff])) < TOL_NEAR_TIE
    assert diff.mean() < 0.02
    ref_ctx.close(); sh.close()


def test_million_row_call_is_chunked_and_counts_every_user_once(torch_cuda):
    """c4's user count through ONE native call (8 row chunks of 131,072 inside skr_eval_fused) on a small catalogue: size-
    independent properties -- every user is evaluated exactly once (Recall@K of a user whose only test item is its own best
    item is 1), the sums equal the sum of per-user rows, and a second call reproduces them bit for bit."""
    from skrec_b200 import _native
    torch = torch_cuda
    U, I, d, K = 1_000_000, 4096, 32, 10
    g = torch.Generator(device="cuda").manual_seed(3)
    ue = torch.randn((U, d), generator=g, device="cuda")
    ie = torch.randn((I, d), generator=g, device="cuda")
    best = torch.empty(U, dtype=torch.int64, device="cuda")
    for u0 in range(0, U, 65536):
        best[u0:u0 + 65536] = (ue[u0:u0 + 65536].double() @ ie.double().T).argmax(1)
    c = _native.Context(0)
    c.set_train_csr(None, None, I)
    c.set_test_csr(np.arange(U + 1, dtype=np.int64), best.cpu().numpy().astype(np.int32), I)
    out = []
    for _ in range(2):
        per = torch.empty((U, 2 * K), dtype=torch.float32, device="cuda")
        sums = torch.zeros(2 * K, dtype=torch.float64, device="cuda")
        c.eval_fused(ue, ie, None, 0, [2, 5], K, precision="tf32r", per_user=per, sums=sums)
        torch.cuda.synchronize()
        out.append((per, sums))
    per, sums = out[0]
    assert torch.equal(per, out[1][0]) and torch.equal(sums, out[1][1])
    assert float(per[:, K - 1].min()) == 1.0                                       # Recall@10 of EVERY user: nobody was skipped
    assert float(per[:, 0].double().mean()) > 0.99999 and float(per[:, K:].min()) >= 0.5   # first, bar float64-vs-float32 near-ties
    assert float((sums - per.double().sum(0)).abs().max()) < 1e-6 and abs(float(sums[0]) - U) < 1e-6
    c.close()
