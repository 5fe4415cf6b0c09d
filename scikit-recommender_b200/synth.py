"""Deterministic synthetic workloads of the shapes BASELINE.json names.

Embeddings are N(0, 0.1^2); user degrees are log-normal; items are drawn with a Zipf-like
popularity without replacement; half of each user's test items are *planted* in the user's own
unmasked top-200 (by exact score) so that the metrics are O(0.1) instead of ~0 -- a parity check
on all-zero metrics would be vacuous.  train and test are disjoint per user; arrays are int32 and
unsorted, like `ImplicitFeedback.to_user_dict()` (dataset.py:150-156) would give.
"""
from collections import OrderedDict

import numpy as np

METRIC_IDS = {"Precision": 1, "Recall": 2, "MAP": 3, "NDCG": 4, "MRR": 5}

# BASELINE.json `configs`, SURVEY.md section 8(d)
CONFIGS = {
    "c1": dict(users=6040, items=3706, d=64, bias=True, nnz_train=800_000, nnz_test=200_000,
               top_k=[10, 20], metric=["Precision", "Recall", "NDCG"], seed=2021,
               name="BPRMF / MovieLens-1M shape"),
    "c2": dict(users=29858, items=40981, d=64, bias=False, nnz_train=810_128, nnz_test=217_242,
               top_k=[20, 50], metric=["Precision", "Recall", "NDCG"], seed=2022,
               name="LightGCN / Gowalla shape"),
    "c3a": dict(users=31668, items=38048, d=64, bias=False, nnz_train=1_237_259, nnz_test=324_147,
                top_k=[20, 50], metric=["Precision", "Recall", "NDCG"], seed=2023,
                name="LightGCN / Yelp2018 shape"),
    "c3b": dict(users=52643, items=91599, d=64, bias=False, nnz_train=2_380_730, nnz_test=603_378,
                top_k=[20, 50], metric=["Precision", "Recall", "NDCG"], seed=2024,
                name="LightGCN / Amazon-Book shape"),
    "c4": dict(users=1_000_000, items=1_000_000, d=128, bias=True, nnz_train=50_000_000, nnz_test=10_000_000,
               top_k=[10, 20, 50, 100], metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], seed=2025,
               name="MultVAE/SelfCF dense scorer, 1M x 1M"),
    "c5": dict(users=2_000_000, items=10_000_000, d=128, bias=False, nnz_train=100_000_000, nnz_test=20_000_000,
               top_k=[100], metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], seed=2026,
               name="item-sharded 10M-item catalogue"),
}


def _degrees(g, n_users, total, lo, hi):
    raw = g.lognormal(0.0, 1.0, n_users)
    deg = np.clip(np.rint(raw * (total / raw.sum())), lo, hi).astype(np.int64)
    # nudge towards the requested total without leaving [lo, hi]
    for _ in range(8):
        diff = int(total - deg.sum())
        if diff == 0:
            break
        room = (deg < hi) if diff > 0 else (deg > lo)
        idx = np.flatnonzero(room)
        if idx.size == 0:
            break
        take = g.choice(idx, size=min(abs(diff), idx.size), replace=False)
        deg[take] += 1 if diff > 0 else -1
    return deg


def _draw_distinct(g, cdf, need, n_items):
    """For every user u draw need[u] distinct items ~ popularity; returns (indptr, items) with the
    per-user draw order preserved."""
    n_users = need.size
    indptr = np.zeros(n_users + 1, np.int64)
    np.cumsum(need, out=indptr[1:])
    out = np.empty(indptr[-1], np.int32)
    have = np.zeros(n_users, np.int64)
    seen_keys = np.zeros(0, np.int64)
    pending = np.arange(n_users)
    while pending.size:
        want = need[pending] - have[pending]
        draws = (want * 1.3).astype(np.int64) + 4
        owner = np.repeat(pending, draws)
        item = np.minimum(np.searchsorted(cdf, g.random(owner.size), side="right"), n_items - 1).astype(np.int64)
        key = owner * n_items + item
        # first occurrence of every (user, item), not already taken in an earlier round
        _, first = np.unique(key, return_index=True)
        first.sort()
        key = key[first]
        if seen_keys.size:
            fresh = ~np.isin(key, seen_keys)
            key = key[fresh]
        owner = key // n_items
        order = np.argsort(owner, kind="stable")
        key, owner = key[order], owner[order]
        start = np.searchsorted(owner, pending, side="left")
        stop = np.searchsorted(owner, pending, side="right")
        rank = np.arange(owner.size) - np.repeat(start, stop - start)
        room = np.repeat(need[pending] - have[pending], stop - start)
        keep = rank < room
        key, owner, rank = key[keep], owner[keep], rank[keep]
        out[indptr[owner] + have[owner] + rank] = (key % n_items).astype(np.int32)
        np.add.at(have, owner, 1)
        seen_keys = np.concatenate([seen_keys, key])
        pending = pending[have[pending] < need[pending]]
    return indptr, out


def heavy_tailed(item_emb, bias, seed):
    """What trained tables look like rather than i.i.d. N(0, 0.1^2): item norms log-normal (sigma 0.5), a handful of
    items (0.05 %, at least 3) at ten times the typical norm, and -- when there is a bias -- a bias that grows with
    the norm (popular items), spread over ~0.3.  In place on copies; -> (item_emb, bias)."""
    g = np.random.default_rng(seed + 99)
    n = item_emb.shape[0]
    f = np.exp(0.5 * g.standard_normal(n)).astype(np.float32)
    out = g.choice(n, size=max(3, n // 2000), replace=False)
    f[out] = 10.0
    emb = (item_emb * f[:, None]).astype(np.float32)
    b = None
    if bias is not None:
        b = (bias + 0.1 * np.log(f)).astype(np.float32)
    return emb, b


def make(users, items, d, nnz_train, nnz_test, seed, bias=False, plant_pool=200, device=None, item_seed=None, norms="iid",
         **_unused):
    """Generate one workload.  Returns a dict: user_emb [U,d] f32, item_emb [I,d] f32, bias [I] f32 or
    None, train/test OrderedDict {user: int32 array} (keys ascending), and the CSR forms
    train_indptr/train_indices/test_indptr/test_indices (rows = users 0..U-1).
    norms="heavy": heavy-tailed item norms, a few outliers and a norm-correlated bias (`heavy_tailed`).
    item_seed: draw the item table (and bias) from their own generator -- ranks of a user-sharded run pass
    the same item_seed and different `seed`s: one replicated catalogue, different user slices."""
    import torch

    g = np.random.default_rng(seed)
    U, I = int(users), int(items)
    user_emb = (g.standard_normal((U, d)) * 0.1).astype(np.float32)
    gi = g if item_seed is None else np.random.default_rng(item_seed)
    item_emb = (gi.standard_normal((I, d)) * 0.1).astype(np.float32)
    b = (gi.standard_normal(I) * 0.01).astype(np.float32) if bias else None
    if norms == "heavy":
        item_emb, b = heavy_tailed(item_emb, b, seed if item_seed is None else item_seed)

    deg_train = _degrees(g, U, nnz_train, 1, max(1, I // 4))
    deg_test = _degrees(g, U, nnz_test, 1, max(1, I // 8))
    n_plant = deg_test // 2
    n_rand = deg_test - n_plant
    perm = g.permutation(I)
    p = 1.0 / (perm + 10.0) ** 0.8
    cdf = np.cumsum(p / p.sum())
    ptr, drawn = _draw_distinct(g, cdf, deg_train + n_rand, I)

    # planted half: uniformly from the user's exact-score top-`plant_pool` among items not drawn
    dev = torch.device(device) if device is not None else torch.device("cpu")
    it = torch.from_numpy(item_emb).to(dev)
    bt = torch.from_numpy(b).to(dev) if b is not None else None
    pool = min(plant_pool, I - int((deg_train + n_rand).max()))
    planted = np.empty((U, pool), np.int32)
    chunk = max(1, min(U, (1 << 28) // max(I, 1)))
    for u0 in range(0, U, chunk):
        u1 = min(U, u0 + chunk)
        s = torch.from_numpy(user_emb[u0:u1]).to(dev) @ it.T
        if bt is not None:
            s = s + bt
        rows = np.repeat(np.arange(u1 - u0), np.diff(ptr[u0:u1 + 1]))
        cols = drawn[ptr[u0]:ptr[u1]].astype(np.int64)
        s[torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev)] = float("-inf")
        planted[u0:u1] = torch.topk(s, pool, dim=1).indices.to("cpu", torch.int32).numpy()
    shuffle = np.argsort(g.random((U, pool)), axis=1)
    planted = np.take_along_axis(planted, shuffle, axis=1)
    n_plant = np.minimum(n_plant, pool)

    train, test = OrderedDict(), OrderedDict()
    tr_ptr = np.zeros(U + 1, np.int64)
    te_ptr = np.zeros(U + 1, np.int64)
    np.cumsum(deg_train, out=tr_ptr[1:])
    np.cumsum(n_rand + n_plant, out=te_ptr[1:])
    tr_idx = np.empty(tr_ptr[-1], np.int32)
    te_idx = np.empty(te_ptr[-1], np.int32)
    for u in range(U):
        row = drawn[ptr[u]:ptr[u + 1]]
        tr = row[:deg_train[u]]
        te = np.concatenate([row[deg_train[u]:], planted[u, :n_plant[u]]])
        tr_idx[tr_ptr[u]:tr_ptr[u + 1]] = tr
        te_idx[te_ptr[u]:te_ptr[u + 1]] = te
        train[u] = tr_idx[tr_ptr[u]:tr_ptr[u + 1]]
        test[u] = te_idx[te_ptr[u]:te_ptr[u + 1]]
    return dict(user_emb=user_emb, item_emb=item_emb, bias=b, train=train, test=test,
                train_indptr=tr_ptr, train_indices=tr_idx, test_indptr=te_ptr, test_indices=te_idx,
                users=U, items=I, d=d)


def make_large(users, items, d, nnz_train, nnz_test, seed, bias=False, plant_pool=200, device="cuda", item_seed=None,
               user_range=None, **_unused):
    """The recipe of `make` for the 10^6-user configs, vectorised with torch on `device` (a GPU): no Python loop over users,
    no dicts.  Same distributions -- N(0, 0.1^2) embeddings, log-normal degrees, Zipf-like item popularity without
    replacement per user, train and test disjoint, half of every user's test items planted among the items the user
    scores highest -- but not the same random streams as `make`.  Planting: an item is a candidate when its score
    (one bf16 tensor-core pass, this is data synthesis) exceeds the user's expected top-`plant_pool` cut sigma_u * z,
    sigma_u = 0.1 ||u||, z = Phi^-1(1 - plant_pool / items); planted items are drawn uniformly from the candidates.
    Returns the dict of `make` without `train` / `test` dicts (use RankingEvaluator.from_csr); embeddings come as CUDA
    tensors (`user_emb`, `item_emb`, `bias`)."""
    import math
    import torch

    dev = torch.device(device)
    U, I = int(users), int(items)
    g = torch.Generator(device=dev).manual_seed(int(seed))
    gi = g if item_seed is None else torch.Generator(device=dev).manual_seed(int(item_seed))
    user_emb = torch.randn((U, d), generator=g, device=dev) * 0.1
    item_emb = torch.randn((I, d), generator=gi, device=dev) * 0.1
    b = torch.randn(I, generator=gi, device=dev) * 0.01 if bias else None

    gn = np.random.default_rng(seed)
    deg_train = _degrees(gn, U, nnz_train, 1, max(1, I // 4))
    deg_test = _degrees(gn, U, nnz_test, 1, max(1, I // 8))
    n_plant = deg_test // 2
    n_rand = deg_test - n_plant
    need = torch.from_numpy(deg_train + n_rand).to(dev)
    perm = torch.randperm(I, generator=g, device=dev).double()
    p = 1.0 / (perm + 10.0) ** 0.8
    cdf = torch.cumsum(p / p.sum(), 0)

    # draws: 1.3x + 4 per user, first occurrences of every (user, item), a random `need[u]` of them kept
    draws = (need.double() * 1.3).long() + 4
    owner = torch.repeat_interleave(torch.arange(U, device=dev), draws)
    item = torch.searchsorted(cdf, torch.rand(owner.numel(), generator=g, device=dev, dtype=torch.float64), right=True).clamp_(max=I - 1)
    key = torch.unique(owner * I + item)  # sorted, distinct
    del owner, item
    owner = key // I
    # random order inside every user: sort by (user, random priority)
    order = torch.argsort(owner.double() + torch.rand(key.numel(), generator=g, device=dev, dtype=torch.float64))
    key = key[order]
    owner = owner[order]
    del order
    start = torch.searchsorted(owner, torch.arange(U + 1, device=dev))
    have = torch.minimum(start[1:] - start[:-1], need)
    rank_in = torch.arange(key.numel(), device=dev) - start[owner]
    keep = rank_in < have[owner]
    key, owner, rank_in = key[keep], owner[keep], rank_in[keep]
    drawn_sorted = torch.sort(key).values  # for the "already drawn" test while planting
    # first deg_train[u] of a user's kept draws are train items, the rest random test items (short users: train first)
    dtr = torch.minimum(torch.from_numpy(deg_train).to(dev), have)
    is_train = rank_in < dtr[owner]
    tr_idx = (key[is_train] % I).int()
    tr_ptr = torch.zeros(U + 1, dtype=torch.int64, device=dev)
    tr_ptr[1:] = torch.cumsum(dtr, 0)
    te_rand_owner = owner[~is_train]
    te_rand_item = (key[~is_train] % I).int()
    del key, owner, rank_in, keep, is_train

    # planted half
    z = math.sqrt(2.0) * float(torch.erfinv(torch.tensor(1.0 - 2.0 * min(plant_pool, I // 4) / I, dtype=torch.float64)))
    sig = 0.1 * user_emb.norm(dim=1)
    it16 = item_emb.to(torch.bfloat16)
    npl = torch.from_numpy(n_plant).to(dev)
    pl_owner, pl_item = [], []
    b16 = None if b is None else b.to(torch.bfloat16)
    chunk = max(1, min(U, (16 << 30) // max(2 * I, 1)))  # 16 GB of bf16 scores per step: few, large steps
    for u0 in range(0, U, chunk):
        u1 = min(U, u0 + chunk)
        sc = user_emb[u0:u1].to(torch.bfloat16) @ it16.T
        if b16 is not None:
            sc += b16
        hit = (sc > (sig[u0:u1] * z).to(torch.bfloat16).unsqueeze(1)).nonzero()  # compared in bf16: a third of the traffic
        del sc
        ck = (hit[:, 0] + u0) * I + hit[:, 1]
        pos = torch.searchsorted(drawn_sorted, ck).clamp_(max=drawn_sorted.numel() - 1)
        ck = ck[drawn_sorted[pos] != ck]  # not among the user's drawn items
        co = ck // I
        order = torch.argsort(co.double() + torch.rand(ck.numel(), generator=g, device=dev, dtype=torch.float64))
        ck, co = ck[order], co[order]
        st = torch.searchsorted(co, torch.arange(u0, u1 + 1, device=dev))
        r_in = torch.arange(ck.numel(), device=dev) - st[co - u0]
        keep = r_in < npl[co]
        pl_owner.append(co[keep])
        pl_item.append((ck[keep] % I).int())
    pl_owner, pl_item = torch.cat(pl_owner), torch.cat(pl_item)
    te_owner = torch.cat([te_rand_owner, pl_owner])
    te_item = torch.cat([te_rand_item, pl_item])
    order = torch.argsort(te_owner, stable=True)  # random items first, then the planted ones, like `make`
    te_owner, te_idx = te_owner[order], te_item[order]
    te_ptr = torch.searchsorted(te_owner, torch.arange(U + 1, device=dev))
    # users whose draws all went to train keep one test item: their best-scoring candidate is not guaranteed, so give them item 0..
    empty = (te_ptr[1:] - te_ptr[:-1]) == 0
    if bool(empty.any()):  # (does not happen with the recipe's degrees: n_rand >= 1 and draws exceed the need)
        raise RuntimeError("synthetic workload: a user without test items")
    return dict(user_emb=user_emb, item_emb=item_emb, bias=b,
                train_indptr=tr_ptr.cpu().numpy(), train_indices=tr_idx.cpu().numpy(),
                test_indptr=te_ptr.cpu().numpy(), test_indices=te_idx.cpu().numpy(), users=U, items=I, d=d)


def make_config(name, scale=1.0, device=None):
    """Workload of a named config; `scale` < 1 shrinks users, items and interactions together."""
    cfg = dict(CONFIGS[name])
    if scale != 1.0:
        cfg["users"] = max(256, int(cfg["users"] * scale))
        cfg["items"] = max(512, int(cfg["items"] * scale))
        cfg["nnz_train"] = max(cfg["users"], int(cfg["nnz_train"] * scale * scale))
        cfg["nnz_test"] = max(cfg["users"], int(cfg["nnz_test"] * scale * scale))
    data = make(device=device, **cfg)
    data["config"] = cfg
    return data


class PredictOnlyModel(object):
    """Stand-in for a trained dot-product recommender that only speaks the reference protocol
    (base.py:73-74): `predict(users)` -> host float32 [B, I]."""

    def __init__(self, user_emb, item_emb, bias=None):
        self.user_emb, self.item_emb, self.bias = user_emb, item_emb, bias

    def predict(self, users):
        import torch
        u = torch.as_tensor(self.user_emb)[torch.as_tensor(np.asarray(users, dtype=np.int64))]
        s = u @ torch.as_tensor(self.item_emb).T
        if self.bias is not None:
            s = s + torch.as_tensor(self.bias)
        return s.cpu().detach().numpy()


class EmbeddingModel(PredictOnlyModel):
    """Same model, additionally offering the fused protocol `eval_embeddings(users) ->
    (user_vecs, item_vecs, bias)`.  Tables may be numpy (host) or torch tensors (device)."""

    def eval_embeddings(self, users):
        idx = np.asarray(users, dtype=np.int64)
        if isinstance(self.user_emb, np.ndarray):
            u = self.user_emb[idx]
        else:
            import torch
            u = self.user_emb[torch.as_tensor(idx, device=self.user_emb.device)]
        return u, self.item_emb, self.bias
