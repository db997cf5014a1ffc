import numpy as _np


class ShapedArray:
  def __init__(self, shape, dtype, weak_type=False):
    self.shape = tuple(shape)
    self.dtype = _np.dtype(dtype)
    self.ndim = len(self.shape)


class Tracer:
  pass


def eval_context():
  import contextlib
  return contextlib.nullcontext()
