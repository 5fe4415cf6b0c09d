"""Generate tests/golden/user_groups.json from the UNMODIFIED reference function.

`skrec.io.dataset` cannot be imported on Python >= 3.10, so the two definitions this needs --
`UserGroup` and `group_users_by_interactions` (skrec/io/dataset.py:698-765) -- are cut out of the reference file with
`ast` and executed as they are, against a stand-in dataset whose `train_data.to_user_dict()` returns the case's dict.
Run in the build container, where /root/reference exists:   python tests/golden/make_groups_golden.py
"""
import ast
import json
import os
from collections import OrderedDict, defaultdict
from typing import List

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/skrec/io/dataset.py"


def reference_function():
    src = open(REF).read()
    tree = ast.parse(src)
    want = [n for n in tree.body if getattr(n, "name", None) in ("UserGroup", "group_users_by_interactions")]
    assert len(want) == 2
    ns = {"np": np, "defaultdict": defaultdict, "List": List, "RSDataset": object}
    exec(compile(ast.Module(body=want, type_ignores=[]), REF, "exec"), ns)
    return ns["group_users_by_interactions"]


class _Train(object):
    def __init__(self, d):
        self._d = d

    def to_user_dict(self):
        return self._d


class _Dataset(object):
    def __init__(self, d):
        self.train_data = _Train(d)


def make_case(seed, n_users, sigma, skip_every=0):
    g = np.random.default_rng(seed)
    deg = np.maximum(1, np.rint(g.lognormal(2.0, sigma, n_users))).astype(np.int64)
    d = OrderedDict()
    for u in range(n_users):
        if skip_every and u % skip_every == 3:
            continue  # users without training data are not in the dict
        d[u] = np.arange(deg[u], dtype=np.int32)
    return d


CASES = [dict(seed=1, n_users=500, sigma=1.0, num_groups=4), dict(seed=2, n_users=3000, sigma=1.3, num_groups=4),
         dict(seed=3, n_users=200, sigma=0.7, num_groups=3, skip_every=7), dict(seed=4, n_users=1000, sigma=1.0, num_groups=5),
         dict(seed=5, n_users=64, sigma=1.5, num_groups=2)]


def main():
    fn = reference_function()
    out = []
    for c in CASES:
        d = make_case(c["seed"], c["n_users"], c["sigma"], c.get("skip_every", 0))
        groups = fn(_Dataset(d), num_groups=c["num_groups"])
        out.append(dict(case=c, groups=[dict(label=g.label, num_users=int(g.num_users), 
                                             # the reference hands every group the LIST of all groups' totals (dataset.py:762 passes
                                             # `num_interactions`, not the loop's `n_interactions`)
                                             num_interactions=[int(x) for x in np.asarray(g.num_interactions).ravel()],
                                             users=np.asarray(g.users).tolist(), activities=np.asarray(g.activities).tolist())
                                        for g in groups]))
    with open(os.path.join(HERE, "user_groups.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out), "cases;", [[g["label"] for g in o["groups"]] for o in out])


if __name__ == "__main__":
    main()
