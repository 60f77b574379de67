"""Importable alias of the ``ca-lanczos_b200/`` package directory (a hyphen is not a valid Python
identifier, so ``import ca_lanczos_b200`` resolves here and re-exports the real package)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ca-lanczos_b200")
__path__[:] = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
