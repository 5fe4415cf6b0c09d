"""B200-native full-ranking evaluation path of scikit-recommender.

Public surface = the reference's evaluation API (skrec/utils/py/evaluator.py,
skrec/utils/py/cython/pyx_eval_matrix.pyx): `RankingEvaluator`, `MetricReport`, `EarlyStopping`,
`eval_score_matrix`, plus the sibling native API `top_k` / `arg_top_k` (pyx_sort.pyx) and the score-provider
`adapters`, `group_users_by_interactions` (dataset.py:707-765) for `evaluate_groups`, and the negative sampler
`randint_choice` / `batch_randint_choice` (utils/py/random.py, randint.h).  Everything is computed by the sm_100a kernels in csrc/ behind the C ABI of
include/skrec_b200.h; importing this package does not need a GPU, evaluating does.
"""
from .report import MetricReport, EarlyStopping
from .evaluator import RankingEvaluator
from .eval_matrix import eval_score_matrix
from .sort import top_k, arg_top_k
from . import adapters
from .groups import UserGroup, group_users_by_interactions
from .random import randint_choice, batch_randint_choice

__all__ = ["MetricReport", "RankingEvaluator", "EarlyStopping", "eval_score_matrix", "top_k", "arg_top_k", "adapters",
           "UserGroup", "group_users_by_interactions", "randint_choice", "batch_randint_choice"]
__version__ = "0.1.0"
