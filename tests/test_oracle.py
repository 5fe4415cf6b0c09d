"""The CPU oracle against the reference: known-answer vector, golden fixtures produced by the
compiled reference, and (where oracle/_ref is present) the compiled reference itself."""
import glob
import os

import numpy as np
import pytest

import oracle


def test_known_answer_survey_a7():
    # SURVEY.md App. A.7, hand-checked against the compiled reference
    s = np.array([[0.9, 0.1, 0.8, 0.3, 0.7, 0.2]], np.float32)
    out, top = oracle.eval_scores(s, [0, 2], [2, 3], [1, 2, 3, 4, 5], 3, return_topk=True)
    assert top.tolist() == [[0, 2, 4]]
    exp = np.array([[0, 0.5, 0.33333334], [0, 0.5, 0.5], [0, 0.25, 0.25],
                    [0, 0.38685283, 0.38685283], [0, 0.5, 0.5]], np.float32)
    assert np.array_equal(out.reshape(5, 3), exp)


def _golden(golden_dir, prefix):
    files = sorted(glob.glob(os.path.join(golden_dir, prefix + "*.npz")))
    assert files, "golden fixtures missing"
    return files


def test_golden_matrix_cases(golden_dir):
    for f in _golden(golden_dir, "matrix_"):
        z = np.load(f)
        got = oracle.eval_scores(z["scores"], z["test_indptr"], z["test_indices"], z["metric"], int(z["top_k"]))
        assert np.array_equal(got, z["expected"]), os.path.basename(f)


def test_golden_evaluator_cases(golden_dir):
    for f in _golden(golden_dir, "evaluator_"):
        z = np.load(f)
        bias = z["bias"] if z["bias"].size else None
        s = z["user_emb"][z["users"]] @ z["item_emb"].T
        if bias is not None:
            s = s + bias
        s = np.ascontiguousarray(s, np.float32)
        oracle.mask_rows(s, z["train_indptr"], z["train_indices"])
        K = int(z["top_k"].max())
        per = oracle.eval_scores(s, z["test_indptr"], z["test_indices"], z["metric"], K)
        mean = oracle.mean_f32(per).reshape(len(z["metric"]), K)[:, np.sort(z["top_k"]) - 1].ravel()
        # the fixture's scores came from the same numpy matmul, so this is exact
        assert np.array_equal(mean, z["expected_values"]), os.path.basename(f)


def test_tie_policy_is_score_desc_id_asc():
    g = np.random.default_rng(5)
    for _ in range(50):
        n, k = int(g.integers(8, 200)), int(g.integers(1, 8))
        s = g.integers(0, 6, size=n).astype(np.float32)  # heavy ties
        idx, val = oracle.topk(s, k)
        exp = np.argsort(-s, kind="stable")[:k]
        assert idx.tolist() == exp.tolist()
        assert np.array_equal(val, s[exp])


def test_neg_inf_and_nan_rank_last():
    s = np.array([1.0, -np.inf, 3.0, np.nan, 2.0, -np.inf], np.float32)
    idx, _ = oracle.topk(s, 6)
    assert idx.tolist() == [2, 4, 0, 1, 3, 5]


def test_topk_refuses_n_less_than_k():
    with pytest.raises(ValueError):
        oracle.topk(np.zeros(3, np.float32), 5)


def test_mean_f32_is_numpy_mean_axis0():
    g = np.random.default_rng(9)
    a = g.random((5000, 37)).astype(np.float32)
    assert np.array_equal(oracle.mean_f32(a), np.mean(a, axis=0))
    assert np.allclose(oracle.sums_f64(a), a.astype(np.float64).sum(0), rtol=0, atol=1e-9)


def test_mask_rows_matches_fancy_index():
    g = np.random.default_rng(2)
    s = g.random((6, 50)).astype(np.float32)
    ref = s.copy()
    rows = [g.choice(50, size=int(g.integers(0, 9)), replace=False).astype(np.int32) for _ in range(6)]
    for r, it in enumerate(rows):
        if len(it):
            ref[r][it] = -np.inf  # evaluator.py:200
    indptr = np.zeros(7, np.int64)
    np.cumsum([len(t) for t in rows], out=indptr[1:])
    oracle.mask_rows(s, indptr, np.concatenate(rows).astype(np.int32))
    assert np.array_equal(s, ref)


def test_scores_close_to_fp32_matmul():
    g = np.random.default_rng(3)
    u = (g.standard_normal((20, 64)) * 0.1).astype(np.float32)
    v = (g.standard_normal((300, 64)) * 0.1).astype(np.float32)
    b = (g.standard_normal(300) * 0.01).astype(np.float32)
    assert np.max(np.abs(oracle.scores(u, v, b) - (u @ v.T + b))) < 1e-6


# ---- against the compiled reference (present here and on the GPU box via oracle/_ref) ----------
needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")


@needs_ref
def test_restatement_equals_compiled_reference_tie_free():
    g = np.random.default_rng(77)
    for trial in range(60):
        B, N = int(g.integers(1, 20)), int(g.integers(30, 900))
        K = int(g.integers(1, min(N, 60)))
        s = np.stack([(g.permutation(N).astype(np.float32) - N / 3) / np.float32(7 * N) for _ in range(B)])
        sizes = g.integers(0, 15, size=B)
        indptr = np.zeros(B + 1, np.int64)
        np.cumsum(sizes, out=indptr[1:])
        indices = np.concatenate([g.choice(N, size=int(n), replace=False) for n in sizes] + [np.zeros(0, np.int64)]).astype(np.int32)
        metric = list(g.permutation(5)[: int(g.integers(1, 6))] + 1)
        got = oracle.eval_scores(s, indptr, indices, metric, K)
        ref = oracle.ref_evaluate_matrix(s.copy(), indptr, indices, metric, K, threads=2)
        assert np.array_equal(got, ref), "trial %d" % trial


@needs_ref
def test_reference_cython_entry_equals_cpp_shim():
    if not oracle.ref_python_available():
        pytest.skip("Cython build of the reference not present")
    g = np.random.default_rng(4)
    s = np.stack([g.permutation(120).astype(np.float32) for _ in range(9)])
    items = [g.choice(120, size=4, replace=False).astype(np.int32) for _ in range(9)]
    a = oracle.ref_eval_score_matrix(s.copy(), items, [1, 2, 3, 4, 5], 7, 2)
    indptr = np.arange(0, 37, 4).astype(np.int64)
    b = oracle.ref_evaluate_matrix(s, indptr, np.concatenate(items), [1, 2, 3, 4, 5], 7)
    assert np.array_equal(a, b)


@needs_ref
def test_tied_rows_same_metrics_when_ties_do_not_mix_hits():
    # Pop-style integer scores (Pop.py:41-44): the reference's order inside a tie group is a heap
    # artefact, but when a tied group holds only misses (or only hits) the metrics agree.
    g = np.random.default_rng(8)
    N, K = 300, 10
    s = np.repeat(np.arange(N // 3, 0, -1), 3).astype(np.float32)[None, :]  # groups of 3 equal scores
    truth = np.array([0, 1, 2, 9, 10, 11], np.int32)  # whole tie groups
    got = oracle.eval_scores(s, [0, 6], truth, [2, 4], 9)  # K multiple of the group size
    ref = oracle.ref_evaluate_matrix(s.copy(), [0, 6], truth, [2, 4], 9)
    assert np.array_equal(got[:, [8, 17]], ref[:, [8, 17]])
