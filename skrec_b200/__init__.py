"""Import alias: `skrec_b200` is the importable name of the package that lives in the
(hyphenated, hence not importable) directory `scikit-recommender_b200/`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "scikit-recommender_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
