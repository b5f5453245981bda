"""Import shim: the product package lives in `mpc-iris-code_b200/` (named after the reference);
a hyphen is not importable, so this package extends its search path to that directory and
re-exports its API.  `import mpc_iris_code_b200 as iris`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mpc-iris-code_b200")
__path__.insert(0, _real)

with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
