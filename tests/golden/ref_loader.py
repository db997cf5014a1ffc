"""Imports the reference's hot-path modules from /root/reference on the NumPy jax shim.

Build-container only (the GPU box has no /root/reference).  Avoids executing
`neural_tangents/__init__.py` (which pulls in empirical/predict/TF code far
outside the hot path) by registering empty parent packages whose `__path__`
points into the read-only reference tree; the submodules themselves
(`_src/stax/*.py`, `_src/utils/*.py`, `_src/batching.py`) run unmodified.
"""
import os
import sys
import types

REFERENCE_ROOT = '/root/reference'
_SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                     'oracle', 'jax_shim')


def load_reference_stax():
  if not os.path.isdir(REFERENCE_ROOT):
    raise RuntimeError('reference tree not present (expected on the GPU box)')
  if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)
  for name, rel in (('neural_tangents', 'neural_tangents'),
                    ('neural_tangents._src', 'neural_tangents/_src'),
                    ('neural_tangents._src.stax', 'neural_tangents/_src/stax')):
    if name not in sys.modules:
      mod = types.ModuleType(name)
      mod.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
      sys.modules[name] = mod
  import importlib
  linear = importlib.import_module('neural_tangents._src.stax.linear')
  elementwise = importlib.import_module('neural_tangents._src.stax.elementwise')
  combinators = importlib.import_module('neural_tangents._src.stax.combinators')
  branching = importlib.import_module('neural_tangents._src.stax.branching')
  ns = types.SimpleNamespace()
  for m in (linear, elementwise, combinators, branching):
    for k, v in vars(m).items():
      if not k.startswith('_'):
        setattr(ns, k, v)
  ns.Kernel = importlib.import_module('neural_tangents._src.utils.kernel').Kernel
  return ns
