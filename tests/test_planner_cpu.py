"""Work list of the tcgen05 main pass (host C++ in the library, `skr_plan_work_host`): properties that must hold for
every shape, checked without a GPU.  A hole or an overlap in the list would silently drop or double-count candidates."""
import numpy as np
import pytest

from skrec_b200 import _native


def _shapes():
    g = np.random.default_rng(7)
    fixed = [(234, 321), (248, 298), (412, 716), (977, 7813), (2048, 9766), (1, 1), (1, 7), (3, 29), (148, 8), (149, 9), (5000, 2)]
    rand = [(int(g.integers(1, 3000)), int(g.integers(1, 12000))) for _ in range(40)]
    return fixed + rand


@pytest.mark.parametrize("n_sm", [148, 132])
@pytest.mark.parametrize("overhead", [4, 6, 10])
def test_work_list_partitions_every_user_tile(n_sm, overhead):
    for n_rt, n_ct in _shapes():
        items, info = _native.plan_work(n_rt, n_ct, n_sm=n_sm, cta_overhead=overhead)
        rt, t0, n, c = items.T.astype(np.int64)
        assert items.shape[0] >= n_rt and (n > 0).all() and (t0 >= 0).all() and (t0 + n <= n_ct).all()
        # every user tile: its ranges tile [0, n_ct) exactly, chunk indices ascend with the ranges
        order = np.lexsort((t0, rt))
        rt_s, t0_s, n_s, c_s = rt[order], t0[order], n[order], c[order]
        first = np.r_[True, rt_s[1:] != rt_s[:-1]]
        last = np.r_[first[1:], True]
        assert np.array_equal(np.unique(rt_s), np.arange(n_rt))
        assert (t0_s[first] == 0).all() and ((t0_s + n_s)[last] == n_ct).all()
        assert ((t0_s + n_s)[~last] == t0_s[~first]).all()
        assert (c_s[first] == 0).all() and (c_s[~first] > c_s[~last]).all()
        # chunk indices stay below `slots` (4 candidate sub-lists per chunk, at most 32 per row)
        assert 1 <= info["min_slots"] <= info["slots"] <= 8 and info["slots"] - info["min_slots"] <= 1
        assert c.max() < info["slots"]
        per_tile = np.bincount(rt, minlength=n_rt)
        assert per_tile.max() <= info["slots"]
        assert info["mixed"] == int(per_tile.min() != per_tile.max()) or n_ct < info["slots"]
        # largest first (the block scheduler then runs longest-processing-time-first), sizes as reported
        assert (n[:-1] >= n[1:]).all() and n.max() == info["max_tiles"]
        # the simulated makespan can not beat a perfect split of the work, nor lose to one chunk per user tile
        total = int(n.sum()) + overhead * items.shape[0]
        assert n.sum() == n_rt * n_ct and info["makespan"] >= -(-total // n_sm)
        single = -(-n_rt // n_sm) * (n_ct + overhead)
        # (up to 3 % may be given away for more, shorter candidate sub-lists when there are many user tiles: the selection
        # stage gains more than the main pass loses)
        assert info["makespan"] * 100 <= single * 103


def test_c2_plan_fills_whole_waves():
    """c2 (29,858 users x 40,981 items -> 234 x 321 tiles): 740 CTAs = 5 full waves of 148 (the bench's plan)."""
    items, info = _native.plan_work(234, 321, n_sm=148, cta_overhead=10)
    assert items.shape[0] == 740 and info["slots"] == 4 and info["min_slots"] == 3 and info["mixed"] == 1


def test_forced_chunks_and_bad_arguments():
    items, info = _native.plan_work(10, 100, chunks=5)
    assert items.shape[0] == 50 and info["slots"] == 5 and info["mixed"] == 0
    items, info = _native.plan_work(10, 3, chunks=8)  # never more chunks than item tiles
    assert items.shape[0] == 30 and info["slots"] == 3
    with pytest.raises(_native.NativeError):
        _native.plan_work(0, 5)
    with pytest.raises(_native.NativeError):
        _native.plan_work(5, 5, n_sm=0)
