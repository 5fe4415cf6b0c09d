"""Host logic of the rows next to the hot path: score-provider adapters (algebra against the
reference models' scoring expressions), CSR ingestion helpers, the lazy {user: items} view.  No GPU."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from skrec_b200 import adapters
from skrec_b200 import evaluator as ev


def _g(seed=0):
    return np.random.default_rng(seed)


def test_dot_product_and_bias_match_bprmf_expression():
    g = _g()
    U, I, b = g.standard_normal((7, 8)).astype(np.float32), g.standard_normal((11, 8)).astype(np.float32), g.standard_normal(11).astype(np.float32)
    sc = adapters.dot_product(U, I, b)
    users = [3, 0, 6]
    uv, iv, bias = sc.eval_embeddings(users)
    assert torch.equal(uv, torch.from_numpy(U[users])) and iv.shape == (11, 8)
    want = U[users] @ I.T + b  # BPRMF.py:84-88
    assert np.allclose(sc.predict(users), want, atol=1e-6)
    assert np.allclose((uv @ iv.T + bias).numpy(), want, atol=1e-6)


def test_two_tower_sum_equals_selfcf_expression():
    g = _g(1)
    u_on, u_tg = g.standard_normal((5, 4)).astype(np.float32), g.standard_normal((5, 4)).astype(np.float32)
    i_on, i_tg = g.standard_normal((9, 4)).astype(np.float32), g.standard_normal((9, 4)).astype(np.float32)
    sc = adapters.two_tower_sum(u_on, i_tg, u_tg, i_on)
    want = u_on @ i_tg.T + u_tg @ i_on.T  # SelfCF.py:235-241
    assert np.allclose(sc.predict(list(range(5))), want, atol=1e-5)
    assert sc.eval_embeddings([1, 2])[0].shape == (2, 8)


def test_neg_euclidean_is_rank_equivalent_to_cml_expression():
    g = _g(2)
    U, I = g.standard_normal((6, 5)).astype(np.float32), g.standard_normal((40, 5)).astype(np.float32)
    sc = adapters.neg_euclidean(U, I)
    got = sc.predict(list(range(6)))
    want = -np.linalg.norm(U[:, None, :] - I[None, :, :], axis=-1)  # CML.py:152
    assert np.array_equal(np.argsort(-got, axis=1, kind="stable"), np.argsort(-want, axis=1, kind="stable"))


def test_decoder_layer_and_item_scores_and_user_index():
    g = _g(3)
    H, W, c = g.standard_normal((4, 6)).astype(np.float32), g.standard_normal((10, 6)).astype(np.float32), g.standard_normal(10).astype(np.float32)
    sc = adapters.decoder_layer(H, W, c, user_index={42: 0, 7: 1, 9: 2, 1: 3})
    assert np.allclose(sc.predict([9, 42]), H[[2, 0]] @ W.T + c, atol=1e-6)  # MultVAE.py:138-141
    pop = adapters.item_scores(np.arange(10, dtype=np.float32), num_users=3)
    assert np.array_equal(pop.predict([0, 2]), np.tile(np.arange(10, dtype=np.float32), (2, 1)))  # Pop.py:41-44


def test_item_shard_protocol_slices_rows():
    from skrec_b200 import dist
    g = _g(4)
    U, I, b = g.standard_normal((3, 4)).astype(np.float32), g.standard_normal((10, 4)).astype(np.float32), g.standard_normal(10).astype(np.float32)
    sc = adapters.dot_product(U, I, b)
    uv, rows, bias, n = sc.eval_embeddings([0, 1, 2], item_shard=(1, 3))
    lo, hi = dist.shard_range(10, 1, 3)
    assert n == 10 and torch.equal(rows, torch.from_numpy(I[lo:hi])) and torch.equal(bias, torch.from_numpy(b[lo:hi]))


def test_csr_rows_and_lazy_view():
    indptr = np.array([0, 2, 2, 5, 6], np.int64)
    indices = np.array([4, 1, 7, 3, 9, 0], np.int32)
    p, i = ev._csr_rows(indptr, indices, [2, 0, 3])
    assert p.tolist() == [0, 3, 5, 6] and i.tolist() == [7, 3, 9, 4, 1, 0]
    p, i = ev._csr_rows(indptr, indices, [2, 0, 3], (3, 8))
    assert p.tolist() == [0, 2, 3, 3] and i.tolist() == [4, 0, 1]
    p, i = ev._csr_rows(indptr, indices, [1])
    assert p.tolist() == [0, 0] and i.size == 0
    lazy = ev._LazyRows(indptr, indices, np.array([0, 2, 3]))
    assert len(lazy) == 3 and 1 not in lazy and 2 in lazy and 17 not in lazy
    assert lazy[2].tolist() == [7, 3, 9] and list(lazy.keys()) == [0, 2, 3]
    with pytest.raises(KeyError):
        lazy[1]
    every = ev._LazyRows(indptr, indices)
    assert len(every) == 4 and 1 in every and every[1].size == 0


def test_from_csr_builds_the_same_evaluator_state_without_a_gpu():
    import scipy.sparse as sp
    g = _g(5)
    tr = sp.random(30, 50, density=0.1, random_state=1, format="csr")
    te = sp.random(30, 50, density=0.05, random_state=2, format="csr")
    e = ev.RankingEvaluator.from_csr(tr, te, metric=["NDCG", "Recall"], top_k=[5, 10])
    users = np.flatnonzero(np.diff(te.indptr) > 0).tolist()
    assert e._all_users == users and e.max_top == 10 and e.metrics == [4, 2]
    assert e.metrics_list == ["NDCG@5", "NDCG@10", "Recall@5", "Recall@10"]
    u = users[0]
    assert np.array_equal(e.user_pos_test[u], te.indices[te.indptr[u]:te.indptr[u + 1]])
    assert np.array_equal(e.user_pos_train[u], tr.indices[tr.indptr[u]:tr.indptr[u + 1]])
    with pytest.raises(RuntimeError):  # no CPU fallback
        if not torch.cuda.is_available():
            e.evaluate(adapters.dot_product(np.zeros((30, 4), np.float32), np.zeros((50, 4), np.float32)))
        else:
            raise RuntimeError("gpu present")


def _pairs_case(seed, n_users=40, n_items=60, n_train=400, n_test=90):
    g = _g(seed)
    def draw(n):
        p = np.unique(np.stack([g.integers(0, n_users, n), g.integers(0, n_items, n)], 1), axis=0)
        return p[g.permutation(p.shape[0])]  # file order is not sorted by user
    return draw(n_train), draw(n_test)


def test_from_pairs_equals_the_reference_dict_construction():
    """`to_user_dict()` restated (dataset.py:148-156: groupby user ascending, items in file order) against from_pairs."""
    import pandas as pd
    tr, te = _pairs_case(11)
    def to_user_dict(p):
        df = pd.DataFrame(p, columns=["user", "item"])
        return OrderedDict((u, d["item"].to_numpy(dtype=np.int32)) for u, d in df.groupby("user"))
    dtr, dte = to_user_dict(tr), to_user_dict(te)
    a = ev.RankingEvaluator(dtr, dte, metric=["Recall", "NDCG"], top_k=[5, 20])
    b = ev.RankingEvaluator.from_pairs(tr, te, metric=["Recall", "NDCG"], top_k=[5, 20])
    assert b._all_users == list(dte.keys()) == a._all_users
    for u in dte:
        assert np.array_equal(b.user_pos_test[u], dte[u])
    for u in dtr:
        assert np.array_equal(b.user_pos_train[u], dtr[u])
    assert all(b.user_pos_train[u].size == 0 for u in range(40) if u not in dtr)
    # the CSRs the native context would be given are identical row for row
    users = a._all_users
    for d_dict, d_lazy in ((dte, b.user_pos_test), (dtr, b.user_pos_train)):
        p0, i0 = ev._rows_to_csr(users, d_dict)
        p1, i1 = ev._rows_to_csr(users, d_lazy)
        assert np.array_equal(p0, p1) and np.array_equal(i0, i1)
        p0, i0 = ev._rows_to_csr(users, d_dict, (10, 35))  # an item shard's column partition
        p1, i1 = ev._rows_to_csr(users, d_lazy, (10, 35))
        assert np.array_equal(p0, p1) and np.array_equal(i0, i1)
    assert a.metrics_list == b.metrics_list and a.max_top == b.max_top
    # sizes: inferred like dataset.py:407-411, or given
    c = ev.RankingEvaluator.from_pairs(None, te, num_users=100, num_items=200)
    assert len(c.user_pos_train) == 0 and c._csr[1][0].size == 101
    with pytest.raises(ValueError):
        ev.RankingEvaluator.from_pairs(tr, te, num_users=5)
    with pytest.raises(ValueError):
        ev.RankingEvaluator.from_pairs(tr, np.array([[0, -1]]))


def test_from_files_reads_the_reference_interaction_files(tmp_path):
    tr, te = _pairs_case(12)
    g = _g(3)
    # UIRT columns, tab separated, no header (dataset.py:27-30, 388-395); only user and item matter
    def write(path, p):
        with open(path, "w") as f:
            for u, i in p:
                f.write("%d\t%d\t%.1f\t%d\n" % (u, i, 1.0, int(g.integers(1, 10 ** 9))))
    write(tmp_path / "toy.train", tr)
    write(tmp_path / "toy.test", te)
    a = ev.RankingEvaluator.from_pairs(tr, te, top_k=10)
    b = ev.RankingEvaluator.from_files(str(tmp_path / "toy.train"), str(tmp_path / "toy.test"), top_k=10)
    for x, y in zip(a._csr[0] + a._csr[1], b._csr[0] + b._csr[1]):
        assert np.array_equal(x, y)
    assert a._all_users == b._all_users
    c = ev.RankingEvaluator.from_files(None, str(tmp_path / "toy.test"), top_k=10)
    assert len(c.user_pos_train) == 0
    with pytest.raises(FileNotFoundError):
        ev.RankingEvaluator.from_files(str(tmp_path / "missing.train"), str(tmp_path / "toy.test"))
    # comma separated UI file
    with open(tmp_path / "toy_ui.test", "w") as f:
        for u, i in te:
            f.write("%d,%d\n" % (u, i))
    d = ev.RankingEvaluator.from_files(None, str(tmp_path / "toy_ui.test"), sep=",", top_k=10)
    assert np.array_equal(d._csr[1][1], b._csr[1][1])


# ---- user activity groups (dataset.py:707-765), the caller above evaluate_groups ------------------------------
def _groups_golden():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "user_groups.json")))


def _groups_case_dict(c):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("mk_groups", os.path.join(os.path.dirname(__file__), "golden", "make_groups_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m, m.make_case(c["seed"], c["n_users"], c["sigma"], c.get("skip_every", 0))


def _same_groups(got, want):
    assert [g.label for g in got] == [w["label"] for w in want]
    for gi, (g, w) in enumerate(zip(got, want)):
        assert g.num_users == w["num_users"] and np.array_equal(g.users, np.array(w["users"]))
        assert np.array_equal(g.activities, np.array(w["activities"]))
        assert int(g.num_interactions) == w["num_interactions"][gi]  # the reference stores the list of all groups' totals


def test_user_groups_equal_the_reference_function_golden():
    from skrec_b200 import groups
    for entry in _groups_golden():
        c = entry["case"]
        _, d = _groups_case_dict(c)
        _same_groups(groups.group_users_by_interactions(d, num_groups=c["num_groups"]), entry["groups"])
        # the same interactions as a CSR pair / scipy matrix / the evaluator's lazy view
        n_users = c["n_users"]
        cnt = np.array([len(d[u]) if u in d else 0 for u in range(n_users)], np.int64)
        indptr = np.r_[0, np.cumsum(cnt)].astype(np.int64)
        indices = np.concatenate([d[u] for u in d]).astype(np.int32)
        _same_groups(groups.group_users_by_interactions((indptr, indices), num_groups=c["num_groups"]), entry["groups"])
        _same_groups(groups.group_users_by_interactions(ev._LazyRows(indptr, indices), num_groups=c["num_groups"]), entry["groups"])
        import scipy.sparse as sp
        m = sp.csr_matrix((np.ones(indices.size, np.float32), indices, indptr), shape=(n_users, int(indices.max()) + 1))
        _same_groups(groups.group_users_by_interactions(m, num_groups=c["num_groups"]), entry["groups"])


def test_user_groups_against_the_reference_source_live():
    import os
    if not os.path.exists("/root/reference/skrec/io/dataset.py"):
        pytest.skip("reference sources are only present in the build container")
    from skrec_b200 import groups
    m, _ = _groups_case_dict(dict(seed=0, n_users=8, sigma=1.0))
    fn = m.reference_function()
    for seed in range(20, 32):
        g = _g(seed)
        c = dict(seed=seed, n_users=int(g.integers(50, 2000)), sigma=float(g.uniform(0.5, 1.6)), num_groups=int(g.integers(2, 6)),
                 skip_every=int(g.choice([0, 5, 9])))
        d = m.make_case(c["seed"], c["n_users"], c["sigma"], c["skip_every"])
        want = fn(m._Dataset(d), num_groups=c["num_groups"])
        got = groups.group_users_by_interactions(d, num_groups=c["num_groups"])
        assert [x.label for x in got] == [x.label for x in want]
        for a, b in zip(got, want):
            assert np.array_equal(a.users, b.users) and np.array_equal(a.activities, b.activities) and a.num_users == b.num_users


def test_user_group_objects_feed_the_evaluator():
    from skrec_b200 import groups
    entry = _groups_golden()[0]
    _, d = _groups_case_dict(entry["case"])
    gs = groups.group_users_by_interactions(d)
    assert sum(len(g) for g in gs) == len(d) and sorted(u for g in gs for u in g) == sorted(d.keys())
    assert list(gs[0])[:3] == gs[0].users[:3].tolist()
    with pytest.raises(IndexError):  # two activity levels cannot make four groups (the reference fails the same way)
        groups.group_users_by_interactions({0: np.arange(3), 1: np.arange(3), 2: np.arange(5)}, num_groups=4)


def test_summed_query_equals_hgn_expression():
    g = _g(21)
    B, L, d, n = 9, 5, 16, 40
    u, union = g.standard_normal((B, d)).astype(np.float32), g.standard_normal((B, d)).astype(np.float32)
    e = g.standard_normal((B, L, d)).astype(np.float32)
    W2, b2 = g.standard_normal((n, d)).astype(np.float32), g.standard_normal(n).astype(np.float32)
    tu, tun, te, tw, tb = (torch.from_numpy(x) for x in (u, union, e, W2, b2))
    # HGN.py:147-163
    res = tu.mm(tw.T) + tb
    res += tun.mm(tw.T)
    res += torch.matmul(te, tw.T.unsqueeze(dim=0)).sum(dim=1)
    sc = adapters.summed_query([u, union, e.sum(1)], W2, b2)
    uv, iv, bias = sc.eval_embeddings(list(range(B)))
    assert np.allclose((uv @ iv.T + bias).numpy(), res.numpy(), atol=1e-4)
    assert np.allclose(sc.predict([3, 1]), res.numpy()[[3, 1]], atol=1e-4)
    assert sc.note == "summed_query" and "summed_query" in adapters.__all__


def test_monotone_activations_leave_every_rank_list_unchanged():
    """SLMRec (sigmoid, SLMRec.py:366-370) and GRU4Rec (final activation, GRU4Rec.py:157-158): dropping a strictly
    increasing activation changes scores, not the order of any user's items."""
    g = _g(22)
    U, I, b = g.standard_normal((6, 8)).astype(np.float32), g.standard_normal((50, 8)).astype(np.float32), g.standard_normal(50).astype(np.float32)
    raw = adapters.dot_product(U, I, b).predict(list(range(6)))
    for act in (torch.sigmoid, torch.tanh):
        ref = act(torch.from_numpy(U) @ torch.from_numpy(I).T + torch.from_numpy(b)).double().numpy()
        keep = np.abs(np.diff(np.sort(ref, 1), axis=1)).min() > 0  # no two scores collapse in the activation's float image
        if keep:
            assert np.array_equal(np.argsort(-ref, 1, kind="stable"), np.argsort(-raw.astype(np.float64), 1, kind="stable"))


def test_transrec_scorer_is_the_reference_expression():
    """TransRec.py:86-93: ratings = -l2_distance((u + g + last).unsqueeze(1), I) + b with l2_distance = torch.norm(a - b)
    (utils/torch.py:24-29).  The adapter's own `predict` is that expression; the fused path is checked on the GPU."""
    import torch
    from skrec_b200 import adapters
    g = torch.Generator().manual_seed(3)
    U, I, d = 7, 23, 12
    ue, ie = torch.randn(U, d, generator=g), torch.randn(I, d, generator=g)
    gt, b = torch.randn(1, d, generator=g), torch.randn(I, generator=g)
    last = torch.randint(0, I, (U,), generator=g)
    sc = adapters.transrec(ue, gt, ie, b, last.numpy())
    transed = ue + gt + ie[last]
    ref = -torch.norm(transed.unsqueeze(1) - ie, p=None, dim=-1) + b
    assert sc.score_fn == "neg_l2" and sc.note == "neg_l2_plus_bias"
    assert np.allclose(sc.predict(list(range(U))), ref.numpy(), atol=1e-6)
    uv, iv, bb = sc.eval_embeddings([2, 5])
    assert torch.equal(uv, transed[[2, 5]]) and iv.shape == (I, d) and torch.equal(bb, b)
