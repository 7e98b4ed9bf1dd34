"""Import alias: the product package lives in ``adaface-dev_b200/`` (hyphen, per the repo layout); this shim
makes it importable as ``adaface_dev_b200`` by pointing the package path at that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "adaface-dev_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
