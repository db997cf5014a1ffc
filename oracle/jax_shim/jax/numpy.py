"""`jax.numpy` -> NumPy forwarding with JAX's immutability semantics.

JAX arrays are immutable: `a *= b` rebinds `a` to a new array.  The reference
relies on this (e.g. `ntk *= dot_sigma` on a kernel shared by two `FanOut`
branches, elementwise.py:452-453), so every array produced here is an
`ImmArray`, a `numpy.ndarray` subclass whose augmented assignments are
out-of-place.
"""
import functools as _functools
import numpy as _np

pi = _np.pi
inf = _np.inf
nan = _np.nan
newaxis = None
float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
bool_ = _np.bool_
float_ = _np.float64
dtype = _np.dtype
ndarray = _np.ndarray


class ImmArray(_np.ndarray):
  def __imul__(self, o): return self * o
  def __iadd__(self, o): return self + o
  def __isub__(self, o): return self - o
  def __itruediv__(self, o): return self / o
  def __ipow__(self, o): return self ** o

  def __array_wrap__(self, arr, context=None, return_scalar=False):
    return _np.asarray(arr).view(ImmArray)


def _imm(x):
  if isinstance(x, _np.ndarray):
    return x.view(ImmArray)
  if isinstance(x, _np.generic):
    return _np.asarray(x).view(ImmArray)
  if isinstance(x, tuple):
    return tuple(_imm(v) for v in x)
  if isinstance(x, list):
    return [_imm(v) for v in x]
  return x


def _wrap(f):
  @_functools.wraps(f)
  def g(*a, **k):
    return _imm(f(*a, **k))
  return g


def _ax(axis):
  return tuple(axis) if isinstance(axis, list) else axis


for _name in dir(_np):
  if _name.startswith('_') or _name in globals():
    continue
  _obj = getattr(_np, _name)
  if isinstance(_obj, type):
    globals()[_name] = _obj
  elif callable(_obj):
    globals()[_name] = _wrap(_obj)

linalg = _np.linalg


@_wrap
def array(x, dtype=None):
  return _np.array(x, dtype=dtype)


@_wrap
def asarray(x, dtype=None):
  return _np.asarray(x, dtype=dtype)


@_wrap
def zeros(shape, dtype=None):
  return _np.zeros(shape, dtype or _np.float64)


@_wrap
def ones(shape, dtype=None):
  return _np.ones(shape, dtype or _np.float64)


@_wrap
def mean(x, axis=None, dtype=None, out=None, keepdims=False):
  return _np.mean(x, axis=_ax(axis), dtype=dtype, keepdims=keepdims)


@_wrap
def var(x, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
  return _np.var(x, axis=_ax(axis), dtype=dtype, ddof=ddof, keepdims=keepdims)


@_wrap
def sum(x, axis=None, dtype=None, out=None, keepdims=False):
  return _np.sum(x, axis=_ax(axis), dtype=dtype, keepdims=keepdims)
