"""Host-side mirror of the reference API (no GPU): constructor semantics, attributes, exceptions and
result strings of RankingEvaluator / MetricReport / EarlyStopping (reference evaluator.py)."""
import numpy as np
import pytest

import oracle
from skrec_b200 import EarlyStopping, MetricReport, RankingEvaluator

TRAIN = {0: np.array([1, 2], np.int32), 1: np.array([3], np.int32)}
TEST = {0: np.array([5], np.int32), 2: np.array([7, 8], np.int32)}


def test_defaults_and_attributes():
    ev = RankingEvaluator(TRAIN, TEST)
    assert ev.metrics == [1, 2, 3, 4, 5] and ev.metrics_num == 5
    assert ev.max_top == 50 and ev.top_show.tolist() == list(range(1, 51))
    assert ev.batch_size == 256 and ev.num_thread == 8
    assert ev.user_pos_train is TRAIN and ev.user_pos_test is TEST
    assert len(ev.metrics_list) == 250 and ev.metrics_list[0] == "Precision@1"


def test_metric_argument_forms_and_order():
    assert RankingEvaluator(None, TEST, metric="NDCG", top_k=[20, 10]).metrics_list == ["NDCG@10", "NDCG@20"]
    ev = RankingEvaluator(None, TEST, metric=("MRR", "Recall"), top_k=[5])
    assert ev.metrics == [5, 2] and ev.metrics_list == ["MRR@5", "Recall@5"]
    assert ev.user_pos_train == {}
    with pytest.raises(TypeError):
        RankingEvaluator(None, TEST, metric=3)
    with pytest.raises(AssertionError):
        RankingEvaluator(None, TEST, metric=["HitRatio"])
    with pytest.raises(AssertionError):
        RankingEvaluator(None, {})


def test_reference_keywords_stay_positional_compatible():
    # base.py:27-29 calls with keywords metric=, top_k=, batch_size=, num_thread=
    ev = RankingEvaluator(TRAIN, TEST, metric=["Precision", "Recall", "NDCG"], top_k=[10, 20], batch_size=64, num_thread=4)
    assert (ev.batch_size, ev.num_thread, ev.max_top) == (64, 4, 20)
    with pytest.raises(TypeError):
        RankingEvaluator(TRAIN, TEST, None, 50, 256, 8, 0)  # new options are keyword-only


def test_report_mapping_and_strings():
    rep = MetricReport(["Precision@10", "NDCG@10"], np.array([0.125, 1 / 3], np.float32))
    assert list(rep.metrics()) == ["Precision@10", "NDCG@10"]
    assert rep["NDCG@10"] == np.float32(1 / 3)
    with pytest.raises(KeyError):
        rep["MAP@10"]
    assert rep.values_str == "\x1b[31m0.12500000  \x1b[0m\t\x1b[32m0.33333334  \x1b[0m"
    assert rep.metrics_str == "\x1b[31mPrecision@10\x1b[0m\t\x1b[32mNDCG@10     \x1b[0m"
    with pytest.raises(AssertionError):
        MetricReport(["a"], [1, 2])


def test_strings_match_compiled_reference_modulo_colour():
    if not oracle.ref_python_available():
        pytest.skip("oracle/_ref not built")
    names = ["Precision", "Recall", "MAP", "NDCG", "MRR"]
    ref = oracle.RefRankingEvaluator(TRAIN, TEST, metric=names, top_k=[1, 5, 10])
    ev = RankingEvaluator(TRAIN, TEST, metric=names, top_k=[1, 5, 10])
    strip = lambda s: __import__("re").sub(r"\x1b\[[0-9]+m", "", s)  # the _ref colorama stub emits no codes
    assert ev.metrics_list == ref.metrics_list
    assert strip(ev.metrics_str) == ref.metrics_str
    vals = np.linspace(0, 1, 15).astype(np.float32)
    from refpkg.evaluator import MetricReport as RefReport
    assert strip(MetricReport(ev.metrics_list, vals).values_str) == RefReport(ref.metrics_list, vals).values_str


def test_colour_cycle_wraps_after_six():
    rep = MetricReport(list("abcdefg"), [0.0] * 7)
    cells = rep.metrics_str.split("\t")
    assert cells[0][:5] == cells[6][:5] == "\x1b[31m" and cells[5][:5] == "\x1b[36m"


def test_early_stopping_contract():
    def rep(v):
        return MetricReport(["NDCG@10"], [v])
    es = EarlyStopping(patience=2)
    assert es.key_metric == "NDCG@10" and list(es.best_result.metrics()) == ["None"]
    assert es(rep(0.5)) is False
    assert es(rep(0.5)) is False      # equal is not an improvement: counter 1
    assert es(rep(0.6)) is False      # improvement resets
    assert es(rep(0.1)) is False and es(rep(0.2)) is True
    assert es.best_result["NDCG@10"] == 0.6
    never = EarlyStopping(patience=0)
    assert [never(rep(0.1)) for _ in range(5)] == [False] * 5
    wrong = EarlyStopping(metric="NDCG@10")
    assert wrong(MetricReport(["Recall@10"], [0.1])) is False  # first report is stored unread (evaluator.py:225-226)
    with pytest.raises(KeyError):
        wrong(MetricReport(["Recall@10"], [0.2]))


def test_evaluate_requires_predict_and_a_gpu():
    ev = RankingEvaluator(TRAIN, TEST, top_k=[1])
    with pytest.raises(AssertionError):
        ev.evaluate(object())
    import torch
    if not torch.cuda.is_available():
        class M(object):
            def predict(self, users):
                return np.zeros((len(users), 10), np.float32)
        with pytest.raises(RuntimeError):  # no CPU fallback
            ev.evaluate(M())


def test_bench_workload_configs():
    """bench.py: both arms print the same `config` object; the default workload is c4 at full size, strong-scaled; the
    reference arm's sample of a large config keeps the catalogue and scales the interactions with the users."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("skr_bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from skrec_b200 import synth
    c4 = bench.workload_config("c4", "strong")
    assert (c4["users"], c4["items"], c4["d"], c4["scaling"]) == (1_000_000, 1_000_000, 128, "strong")
    assert c4["top_k"] == [10, 20, 50, 100] and len(c4["metrics"]) == 5 and c4["workload"].startswith("c4")
    s = bench.sample_config("c4")
    assert (s["users"], s["items"], s["d"]) == (2048, 1_000_000, 128)
    assert s["nnz_train"] == 102_400 and s["nnz_test"] == 20_480 and s["item_seed"] == synth.CONFIGS["c4"]["seed"] + 7
    assert synth.CONFIGS["c4"]["users"] == 1_000_000  # the table itself is not touched
    assert "c4" in bench.STRONG and "c3b" in bench.STRONG and "c2" not in bench.STRONG
    assert bench._span([5, 6, 7, 8]) == (5, 4) and bench._span([5, 7, 8]) is None and bench._span([]) is None


def test_set_valued_rows_are_accepted_like_the_reference():
    """BERT4Rec hands `eval_score_matrix` Python sets (bert4rec_utils.py:25); the reference converts them through
    Cython's cset[int].  Same for set-valued evaluator dicts."""
    from skrec_b200.evaluator import _as_i32, _dict_to_csr
    assert sorted(_as_i32({3, 1, 2}).tolist()) == [1, 2, 3] and _as_i32({3, 1, 2}).dtype == np.int32
    assert _as_i32([4, 5]).tolist() == [4, 5] and _as_i32(np.array([[7], [8]])).tolist() == [7, 8]
    ptr, idx = _dict_to_csr([0, 1, 2], {0: {9}, 1: frozenset(), 2: [1, 2]})
    assert ptr.tolist() == [0, 1, 1, 3] and idx.tolist() == [9, 1, 2]


def test_sharding_is_opt_in_and_arguments_are_checked_before_work():
    """`evaluate` must not become a collective just because torch.distributed is initialised."""
    ev = RankingEvaluator({0: np.array([1])}, {0: np.array([2])})
    assert ev.shard_users is False
    assert ev._shard(10) == (0, 1, 0, 10)
    users, key = ev._resolve_users([0, 5, 0])
    assert users == [0, 0] and key[0] == "subset" and key[1] == 2
    assert ev._resolve_users(None) == ([0], ("all",))
