"""User activity groups for grouped evaluation -- the caller right above the hot path.

`group_users_by_interactions` gives what the reference's function of the same name gives
(skrec/io/dataset.py:707-765, used by `AbstractRecommender.evaluate_group`, base.py:66-71): users are binned by their
number of training interactions into `num_groups` ranges of roughly equal interaction mass; labels "< a", "[a, b)",
"≥ b".  Here it works on per-user counts with array operations (no dict of lists per activity level) and takes a
{user: items} mapping, a lazy CSR view, a scipy CSR matrix or an (indptr, indices) pair.  Feed the result to
`RankingEvaluator.evaluate_groups(model, groups)`: one evaluation for all groups instead of one per group.
"""
from typing import List

import numpy as np


class UserGroup(object):
    """Same attributes as the reference's UserGroup (dataset.py:698-704).  `num_interactions` is the group's own
    total; the reference stores the list of ALL groups' totals in every group (dataset.py:762 passes the list
    `num_interactions` instead of the loop variable `n_interactions`) and never reads it (base.py:66-71 uses
    `.users` and `.label`)."""

    def __init__(self, users, num_interactions, activities, label):
        self.label = label
        self.num_users = len(users)
        self.num_interactions = num_interactions
        self.users = users
        self.activities = activities

    def __iter__(self):  # so that a group can be handed to evaluate_groups / evaluate(test_users=...) directly
        return iter(self.users.tolist())

    def __len__(self):
        return self.num_users


def _users_and_counts(train):
    """-> (user ids in the order `to_user_dict()` lists them, their interaction counts)"""
    if hasattr(train, "indptr") and not isinstance(train, dict):  # scipy CSR or the evaluator's lazy view
        indptr = np.asarray(train.indptr, dtype=np.int64)
        keys = getattr(train, "_keys", None)
        if keys is not None:
            users = np.asarray(keys, dtype=np.int64)
            return users, indptr[users + 1] - indptr[users]
        cnt = np.diff(indptr)
        users = np.flatnonzero(cnt > 0)  # a dict built by groupby has no empty users (dataset.py:150-155)
        return users, cnt[users]
    if isinstance(train, tuple) and len(train) == 2:
        cnt = np.diff(np.asarray(train[0], dtype=np.int64))
        users = np.flatnonzero(cnt > 0)
        return users, cnt[users]
    users = np.fromiter(train.keys(), dtype=np.int64, count=len(train))
    cnt = np.fromiter((len(v) for v in train.values()), dtype=np.int64, count=len(train))
    return users, cnt


def group_users_by_interactions(train, num_groups=4) -> List[UserGroup]:
    users, cnt = _users_and_counts(train)
    if users.size == 0:
        raise ValueError("no training interactions to group users by")
    levels, n_at_level = np.unique(cnt, return_counts=True)  # activity levels ascending, users per level
    mass = levels * n_at_level                                # interactions per level
    # boundaries: for each of the first num_groups - 1 groups take levels until the group's share of the REMAINING
    # mass is met, choosing the nearer of the two candidate cuts (ties go to the later one), at least one level
    bounds, start = [], 0
    for g in range(num_groups - 1):
        rest = np.cumsum(mass[start:])
        share = rest[-1] / (num_groups - g)
        i = max(int(np.searchsorted(rest, share)), 1)
        take = i if (share - rest[i - 1] < rest[i] - share) else i + 1
        start += take
        bounds.append(start)
        if start >= levels.size:
            raise IndexError("fewer distinct activity levels than the %d groups need" % num_groups)
    cuts = levels[bounds]
    labels = ["< %s" % cuts[0]] + ["[%s, %s)" % (a, b) for a, b in zip(cuts[:-1], cuts[1:])] + ["≥ %s" % cuts[-1]]
    # users of a group: by ascending activity level, inside a level in the order the mapping lists them
    order = np.argsort(cnt, kind="stable")
    level_end = np.cumsum(n_at_level)
    edges = [0] + bounds + [levels.size]
    groups = []
    for gi, label in enumerate(labels):
        l0, l1 = edges[gi], edges[gi + 1]
        u0 = 0 if l0 == 0 else int(level_end[l0 - 1])
        u1 = int(level_end[l1 - 1])
        groups.append(UserGroup(users[order[u0:u1]], mass[l0:l1].sum(), levels[l0:l1], label))
    return groups
