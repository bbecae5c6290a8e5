"""Import alias: the product package lives in the directory `tiny-differentiable-simulator_b200/`
(not a valid Python identifier), this shim makes it importable as `tds_b200`."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "tiny-differentiable-simulator_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
