"""B200-native full-ranking evaluation path of scikit-recommender.

Public surface = the reference's evaluation API (skrec/utils/py/evaluator.py,
skrec/utils/py/cython/pyx_eval_matrix.pyx): `RankingEvaluator`, `MetricReport`, `EarlyStopping`,
`eval_score_matrix`, plus the sibling native API `top_k` / `arg_top_k` (pyx_sort.pyx) and the score-provider
`adapters`.  Everything is computed by the sm_100a kernels in csrc/ behind the C ABI of
include/skrec_b200.h; importing this package does not need a GPU, evaluating does.
"""
from .report import MetricReport, EarlyStopping
from .evaluator import RankingEvaluator
from .eval_matrix import eval_score_matrix
from .sort import top_k, arg_top_k
from . import adapters

__all__ = ["MetricReport", "RankingEvaluator", "EarlyStopping", "eval_score_matrix", "top_k", "arg_top_k", "adapters"]
__version__ = "0.1.0"
