"""Importable alias of the product package.

The package directory is named ``asr-rescoring_b200`` (repo layout contract),
which is not a valid Python identifier; this module redirects
``import asr_rescoring_b200`` to it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "asr-rescoring_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
