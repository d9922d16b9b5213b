"""Importable alias of the ``ngp-encode-server_b200/`` package directory (a hyphen is not
a valid Python identifier).  All code lives there; this only redirects the package path."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ngp-encode-server_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
