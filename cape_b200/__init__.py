"""Import alias: ``import cape_b200`` loads the package that lives in ``category-agnostic-pose-estimation_b200/``
(a directory name that is not a Python identifier)."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "category-agnostic-pose-estimation_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _f
