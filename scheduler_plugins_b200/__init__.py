"""Importable alias of the product package.

The package directory is `scheduler-plugins_b200/` (the name the build contract fixes); a
hyphen cannot appear in a Python module name, so this alias points `__path__` at it and
executes its `__init__`.  Nothing else lives here.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "scheduler-plugins_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
