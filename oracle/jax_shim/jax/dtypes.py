import numpy as _np


def canonicalize_dtype(dtype):
  return _np.dtype(dtype)  # x64 enabled: float64 stays float64
