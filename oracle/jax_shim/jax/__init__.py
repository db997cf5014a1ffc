"""NumPy-backed stand-in for the subset of `jax` the reference hot path uses.

See ../README.md.  Float64 everywhere (the analogue of `jax_enable_x64=True`).
"""
import functools
import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import lax, random, tree_util, core, ops, dtypes, nn, image  # noqa: F401
from . import scipy  # noqa: F401
from . import example_libraries  # noqa: F401

Array = _np.ndarray


class ShapeDtypeStruct:
  def __init__(self, shape, dtype):
    self.shape = tuple(shape)
    self.dtype = _np.dtype(dtype)
    self.ndim = len(self.shape)


def eval_shape(fun, *args, **kwargs):
  """Shape inference by running `fun` on zero arrays (cheap at these sizes)."""
  def is_shaped(x):
    return isinstance(x, (core.ShapedArray, ShapeDtypeStruct))

  def to_zeros(x):
    return _np.zeros(x.shape, x.dtype) if is_shaped(x) else x

  args = tree_util.tree_map(to_zeros, args, is_leaf=is_shaped)
  kwargs = tree_util.tree_map(to_zeros, kwargs, is_leaf=is_shaped)
  out = fun(*args, **kwargs)

  def to_shaped(x):
    x = _np.asarray(x)
    return ShapeDtypeStruct(x.shape, x.dtype)

  return tree_util.tree_map(to_shaped, out)


def jit(fun=None, **kwargs):
  if fun is None:
    return lambda f: f
  return fun


def vmap(fun, in_axes=0, out_axes=0):
  raise NotImplementedError('vmap')


def grad(fun, *a, **k):
  raise NotImplementedError('grad')


def pmap(fun, *a, **k):
  raise NotImplementedError('pmap')


def device_put(x, device=None):
  return x


def devices(backend=None):
  return ['cpu0']


def default_backend():
  return 'cpu'


class custom_jvp:
  """Only the primal function is ever evaluated on this path."""

  def __init__(self, fun, nondiff_argnums=()):
    self.fun = fun
    functools.update_wrapper(self, fun)

  def __call__(self, *args, **kwargs):
    return self.fun(*args, **kwargs)

  def defjvp(self, jvp):
    return jvp


class _Config:
  def update(self, name, value):
    pass

  jax_enable_x64 = True


config = _Config()
