"""Generate tests/golden/*.npz from the UNMODIFIED, compiled reference (oracle/_ref).

Run in the build container, where /root/reference exists:
    python oracle/build_ref.py && python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py) to
outputs of the reference's own code: `eval_score_matrix` (pyx_eval_matrix.pyx:22-37) and
`RankingEvaluator.evaluate` (evaluator.py:163-214).  Score rows are tie-free (a scaled random
permutation), so the reference's rank list is unique.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def tie_free_scores(g, B, N):
    s = np.empty((B, N), np.float32)
    for r in range(B):
        s[r] = (g.permutation(N).astype(np.float32) - N / 2) / np.float32(N)
    return s


def case_matrix(seed, B, N, K, metric, max_truth, with_neg_inf):
    g = np.random.default_rng(seed)
    s = tie_free_scores(g, B, N)
    test_items = []
    for r in range(B):
        n = int(g.integers(0, max_truth + 1)) if r % 7 else 0  # some empty truth sets
        t = g.choice(N, size=n, replace=False).astype(np.int32)
        if n > 2 and r % 5 == 0:
            t = np.concatenate([t, t[:2]])  # duplicates: the reference's set drops them
        test_items.append(t)
    if with_neg_inf:  # one masked (train) item per row, like evaluator.py:200
        for r in range(B):
            s[r, int(g.integers(0, N))] = -np.inf
    out = oracle.ref_eval_score_matrix(s.copy(), test_items, list(metric), K, 4)
    indptr = np.zeros(B + 1, np.int64)
    np.cumsum([len(t) for t in test_items], out=indptr[1:])
    indices = np.concatenate(test_items).astype(np.int32) if indptr[-1] else np.zeros(0, np.int32)
    return dict(scores=s, test_indptr=indptr, test_indices=indices, metric=np.array(metric, np.int32),
                top_k=np.int32(K), expected=out)


def case_evaluator(seed, U, I, d, K_list, metric, bias):
    g = np.random.default_rng(seed)
    ue = (g.standard_normal((U, d)) * 0.1).astype(np.float32)
    ie = (g.standard_normal((I, d)) * 0.1).astype(np.float32)
    b = (g.standard_normal(I) * 0.01).astype(np.float32) if bias else None
    train, test = {}, {}
    for u in range(U):
        n_tr = int(g.integers(0, 30))
        n_te = int(g.integers(0, 12))
        picks = g.choice(I, size=n_tr + n_te, replace=False).astype(np.int32)
        if n_tr and u % 11:
            train[u] = picks[:n_tr]
        if n_te:
            test[u] = picks[n_tr:]
    # plant some test items at the top so metrics are not ~0
    full = ue @ ie.T + (b if b is not None else 0)
    for u in list(test.keys())[::2]:
        s = full[u].copy()
        if u in train:
            s[train[u]] = -np.inf
        top = np.argsort(-s, kind="stable")[:30]
        extra = g.choice(top, size=3, replace=False).astype(np.int32)
        extra = np.array([x for x in extra if u not in train or x not in set(train[u].tolist())], np.int32)
        test[u] = np.unique(np.concatenate([test[u], extra])).astype(np.int32)

    class Model(object):
        def predict(self, users):
            s = ue[np.asarray(users)] @ ie.T
            if b is not None:
                s = s + b
            return np.ascontiguousarray(s, dtype=np.float32)

    names = {1: "Precision", 2: "Recall", 3: "MAP", 4: "NDCG", 5: "MRR"}
    ev = oracle.RefRankingEvaluator(train, test, metric=[names[m] for m in metric], top_k=list(K_list),
                                    batch_size=37, num_thread=3)
    rep = ev.evaluate(Model())
    users = sorted(test.keys())
    tr_ptr, tr_idx = oracle.dicts_to_csr(users, train)
    te_ptr, te_idx = oracle.dicts_to_csr(users, test)
    return dict(user_emb=ue, item_emb=ie, bias=b if b is not None else np.zeros(0, np.float32),
                users=np.array(users, np.int64), train_indptr=tr_ptr, train_indices=tr_idx,
                test_indptr=te_ptr, test_indices=te_idx, metric=np.array(metric, np.int32),
                top_k=np.array(K_list, np.int32), expected_values=np.array(list(rep.values()), np.float32),
                expected_names=np.array(list(rep.metrics())), values_str=np.array(rep.values_str),
                metrics_str=np.array(ev.metrics_str))


def main():
    assert oracle.ref_python_available(), "build the reference first: python oracle/build_ref.py"
    cases = {
        "matrix_all5_k10": case_matrix(11, 48, 257, 10, [1, 2, 3, 4, 5], 9, False),
        "matrix_prn_k50_masked": case_matrix(12, 33, 1500, 50, [1, 2, 4], 20, True),
        "matrix_order_k3": case_matrix(13, 16, 40, 3, [5, 4, 3, 2, 1], 5, False),
        "matrix_k_eq_n": case_matrix(14, 8, 20, 20, [1, 4], 6, False),
        "evaluator_bias": case_evaluator(21, 150, 400, 16, [5, 10, 20], [1, 2, 3, 4, 5], True),
        "evaluator_nobias": case_evaluator(22, 97, 700, 64, [20, 50], [1, 2, 4], False),
    }
    for name, c in cases.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **c)
        print(name, {k: getattr(v, "shape", None) for k, v in c.items()})


if __name__ == "__main__":
    main()
