"""Importable alias of the package directory `neural-tangents_b200/`.

The product lives in `neural-tangents_b200/` (the name the build contract asks
for); a hyphen is not a valid Python identifier, so this stub points the package
search path there and executes its `__init__.py`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'neural-tangents_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
  exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _os, _f
