"""`Kernel` dataclass — layout contract of `_src/utils/kernel.py:27-189`.

Arrays are NumPy (host) arrays.  Spatial covariances use the reference's zipped
layout `nngp[N1,N2,H,H',W,W']`, or `[N1,N2,W,W',H,H']` when `is_reversed`
(`_src/utils/kernel.py:64-69`); the CUDA library always works in the
non-reversed order and `stax.py` converts at the API boundary.
"""
import dataclasses
from typing import Any, Optional, Tuple

import numpy as np


def _reverse_zipped(m: Optional[np.ndarray], start_axis: int) -> Optional[np.ndarray]:
  """`_src/utils/utils.py:465-480`."""
  if m is None or m.ndim <= start_axis:
    return m
  n_pairs = (m.ndim - start_axis) // 2
  perm = tuple(range(start_axis))
  for k in reversed(range(n_pairs)):
    perm += (start_axis + 2 * k, start_axis + 2 * k + 1)
  return np.transpose(m, perm)


@dataclasses.dataclass(frozen=True)
class Kernel:
  """Fields of `_src/utils/kernel.py:124-145`."""
  nngp: np.ndarray
  ntk: Optional[np.ndarray]
  cov1: np.ndarray
  cov2: Optional[np.ndarray]
  x1_is_x2: Any
  is_gaussian: bool
  is_reversed: bool
  is_input: bool
  diagonal_batch: bool
  diagonal_spatial: bool
  shape1: Optional[Tuple[int, ...]]
  shape2: Optional[Tuple[int, ...]]
  batch_axis: int
  channel_axis: int
  mask1: Optional[np.ndarray] = None
  mask2: Optional[np.ndarray] = None

  def replace(self, **kwargs) -> 'Kernel':
    return dataclasses.replace(self, **kwargs)

  def asdict(self):
    return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}

  def astuple(self):
    return tuple(getattr(self, f.name) for f in dataclasses.fields(self))

  def slice(self, n1_slice: slice, n2_slice: slice) -> 'Kernel':
    """`_src/utils/kernel.py:151-167` (used by `batch` on Kernel inputs)."""
    cov1 = self.cov1[n1_slice]
    cov2 = self.cov1[n2_slice] if self.cov2 is None else self.cov2[n2_slice]
    ntk = self.ntk
    return self.replace(
        cov1=cov1,
        nngp=self.nngp[n1_slice, n2_slice],
        cov2=cov2,
        ntk=ntk if ntk is None or ntk.ndim == 0 else ntk[n1_slice, n2_slice],
        shape1=(cov1.shape[0],) + tuple(self.shape1[1:]),
        shape2=(cov2.shape[0],) + tuple(self.shape2[1:]))

  def reverse(self) -> 'Kernel':
    """`_src/utils/kernel.py:169-189`."""
    batch_ndim = 1 if self.diagonal_batch else 2
    return self.replace(cov1=_reverse_zipped(self.cov1, batch_ndim),
                        nngp=_reverse_zipped(self.nngp, 2),
                        cov2=_reverse_zipped(self.cov2, batch_ndim),
                        ntk=_reverse_zipped(self.ntk, 2),
                        is_reversed=not self.is_reversed)

  def __mul__(self, other: float) -> 'Kernel':
    """`_src/utils/kernel.py:426-433`."""
    var = other**2
    return self.replace(cov1=var * self.cov1, nngp=var * self.nngp,
                        cov2=None if self.cov2 is None else var * self.cov2,
                        ntk=None if self.ntk is None else var * self.ntk)

  __rmul__ = __mul__

  def __add__(self, other: float) -> 'Kernel':
    """`_src/utils/kernel.py:435-439`."""
    var = other**2
    return self.replace(cov1=var + self.cov1, nngp=var + self.nngp,
                        cov2=None if self.cov2 is None else var + self.cov2)

  __sub__ = __add__

  def __truediv__(self, other: float) -> 'Kernel':
    return self.__mul__(1. / other)
