"""Importable alias of the ``point-cloud-reid_b200/`` package directory (a hyphen is not a valid
module name).  ``import pcreid_b200.ops`` resolves to ``point-cloud-reid_b200/ops``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "point-cloud-reid_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
