"""Importable alias of the package directory ``p-div-gnn_b200/``.

The repository layout names the package ``p-div-gnn_b200`` (not a valid Python
identifier); this shim makes ``import pdivgnn_b200`` resolve every submodule there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "p-div-gnn_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
