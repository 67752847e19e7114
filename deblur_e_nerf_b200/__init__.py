"""Importable alias of the ``deblur-e-nerf_b200/`` package directory (a hyphen cannot
appear in a Python module name): ``import deblur_e_nerf_b200`` resolves sub-modules from
``../deblur-e-nerf_b200`` and runs that directory's ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "deblur-e-nerf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _fh
