"""Result containers of the evaluation API.

`MetricReport` and `EarlyStopping` keep the observable behaviour of the reference classes
(skrec/utils/py/evaluator.py:17-54 and :217-246): insertion-ordered name -> value mapping,
KeyError on unknown names, the two coloured tab-joined strings that the reference's loggers print
(`f"{v:.8f}".ljust(12)` cells, colours cycling red, green, yellow, blue, magenta, cyan;
evaluator.py:14,25-43), and a strict-improvement patience counter keyed on one metric.
"""
from collections import OrderedDict
from itertools import cycle

try:
    from colorama import Fore as _Fore, Style as _Style
    _PALETTE = (_Fore.RED, _Fore.GREEN, _Fore.YELLOW, _Fore.BLUE, _Fore.MAGENTA, _Fore.CYAN)
    _RESET = _Style.RESET_ALL
except ImportError:  # colorama absent: emit the ANSI sequences it would have produced
    _PALETTE = tuple("\x1b[%dm" % code for code in (31, 32, 33, 34, 35, 36))
    _RESET = "\x1b[0m"

_CELL = 12


def colour_join(cells):
    """Left-justify every cell to 12 characters, colour it, join with tabs."""
    return "\t".join(colour + str(cell).ljust(_CELL) + _RESET for colour, cell in zip(cycle(_PALETTE), cells))


class MetricReport(object):
    def __init__(self, metrics, values):
        n_names, n_values = len(metrics), len(values)
        assert n_names == n_values, f"The lengths of metrics and values are not equal ({n_names}!={n_values})."
        self._results = OrderedDict()
        for name, value in zip(metrics, values):
            self._results[name] = value

    # mapping protocol -------------------------------------------------------------------------
    def metrics(self):
        return self._results.keys()

    def values(self):
        return self._results.values()

    def items(self):
        return self._results.items()

    @property
    def results(self):
        return self._results

    def __getitem__(self, name):
        try:
            return self._results[name]
        except KeyError:
            raise KeyError(name) from None

    def __str__(self):
        return str(self._results)

    # log strings --------------------------------------------------------------------------------
    @property
    def metrics_str(self):
        return colour_join(self._results.keys())

    @property
    def values_str(self):
        return colour_join("%.8f" % v for v in self._results.values())


class EarlyStopping(object):
    """Call with each new MetricReport; returns True once `patience` consecutive reports failed
    to beat the best value of `metric` strictly (patience <= 0 never stops)."""

    def __init__(self, metric="NDCG@10", patience=100):
        self._metric = metric
        self._patience = patience
        self._best_score = None
        self._counter = 0

    @property
    def key_metric(self):
        return self._metric

    @property
    def best_result(self):
        return self._best_score if self._best_score is not None else MetricReport(["None"], [0])

    def __call__(self, val_result):
        best = self._best_score
        if best is not None and val_result[self._metric] <= best[self._metric]:
            self._counter += 1
            return 0 < self._patience <= self._counter
        self._best_score = val_result
        self._counter = 0
        return False
