"""Import alias: ``pcbgpu`` is the package in ``rl-environment-for-component-placement_b200/``.

The project directory name is not a valid Python identifier, so this shim points
the package search path at it and runs its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "rl-environment-for-component-placement_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
