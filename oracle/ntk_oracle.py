"""CPU oracle: NumPy restatement of the reference's analytic NNGP/NTK kernel path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module.
The product (`neural_tangents_b200`) never imports it and has no CPU fallback.

Parity status: PINNED against the reference's own Python code.  The reference
(`/root/reference`, pure Python on JAX) cannot be imported as-is in this
container (no jax wheel), so `oracle/jax_shim` provides a NumPy stand-in for the
handful of `jax.numpy`/`jax.lax` primitives the hot path calls and
`tests/golden/generate_golden.py` executes the *reference's own* `stax`
`kernel_fn` code on top of it in float64.  The resulting vectors are committed
under `tests/golden/` and this oracle is checked against them
(`tests/test_oracle_golden.py`), together with the reference's literal golden
array (`tests/stax/stax_test.py:515-539`) and its exact identities.

Every function cites the reference lines it restates (paths relative to
`/root/reference/neural_tangents/_src/`).

Layout.  The oracle always works in the canonical zipped layout
`[n1, n2, h, h', w, w']` (reference `is_reversed == False`,
`utils/kernel.py:32-36`); the reference's per-`Conv` axis reversal
(`stax/linear.py:1344-1349, 3365-3378`) is tracked only as the boolean
`is_reversed` and applied on export by `to_reference_layout`.

A network is described by a plain-data *spec* tree (nested tuples) so the oracle
shares no code with the product front end:

  ('serial', [spec, ...])
  ('dense', W_std, b_std)
  ('conv', (kh, kw), (sh, sw), 'SAME'|'VALID'|'CIRCULAR', W_std, b_std)
  ('abrelu', a, b, do_stabilize)          # Relu == ('abrelu', 0., 1., False)
  ('erf', a, b, c)
  ('avgpool', (wh, ww), (sh, sw), 'SAME'|'VALID'|'CIRCULAR', normalize_edges)
  ('sumpool', (wh, ww), (sh, sw), padding)   ('gsp',)      # SumPool / GlobalSumPool
  ('gelu',)  ('sin', a, b, c)  ('cos', a, b, c)  ('rbf', gamma)  ('layernorm', eps)
  ('gap',) ('flatten',) ('identity',)
  ('fanout', n) ('parallel', [spec, ...]) ('faninsum',)
"""
from __future__ import annotations

import dataclasses
import math
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------
# State carried between layers (restates utils/kernel.py:124-145, only the
# fields the hot path reads).
# --------------------------------------------------------------------------
@dataclasses.dataclass
class OState:
  nngp: np.ndarray
  ntk: Optional[np.ndarray]          # 0-d array == "scalar zero" (requirements.py:807)
  cov1: np.ndarray
  cov2: Optional[np.ndarray]         # None <=> x2 is None
  is_gaussian: bool
  is_reversed: bool
  spatial: bool                      # True while the 4 spatial axes are present
  shape1: Tuple[int, ...]
  shape2: Tuple[int, ...]

  def replace(self, **kw) -> 'OState':
    return dataclasses.replace(self, **kw)


# --------------------------------------------------------------------------
# Padding arithmetic (lax.padtype_to_pads semantics; SURVEY Appendix A.3,
# used at stax/linear.py:3051-3058 and by lax.reduce_window/conv).
# --------------------------------------------------------------------------
def same_pads(n: int, k: int, s: int) -> Tuple[int, int, int]:
  """Returns (out, lo, hi) for SAME padding."""
  out = -(-n // s)
  tot = max((out - 1) * s + k - n, 0)
  lo = tot // 2
  return out, lo, tot - lo


def valid_out(n: int, k: int, s: int) -> int:
  return (n - k) // s + 1 if n >= k else 0


# --------------------------------------------------------------------------
# Input layer: requirements.py:641-830 (_inputs_to_kernel), 585-638 (_cov),
# 542-553 (_cov_full_batch_full_spatial), 572-582 (_cov_diag_batch).
# --------------------------------------------------------------------------
def inputs_to_state(x1: np.ndarray, x2: Optional[np.ndarray], compute_ntk: bool,
                    dtype=np.float64) -> OState:
  if not isinstance(x1, np.ndarray) or not (x2 is None or isinstance(x2, np.ndarray)):
    raise TypeError('inputs must be ndarrays')                # requirements.py:754-758
  if x1.ndim < 2:
    raise ValueError('Inputs must be at least 2D')            # requirements.py:782-784
  x1 = x1.astype(dtype)                                       # requirements.py:794
  x2c = None if x2 is None else x2.astype(dtype)
  c = x1.shape[-1]
  if x1.ndim == 2:
    # FCN: nngp = x1 x2^T / d, cov = |x|^2 / d
    y = x1 if x2c is None else x2c
    nngp = x1 @ y.T / c
    cov1 = np.sum(x1 * x1, axis=1) / c
    cov2 = None if x2c is None else np.sum(x2c * x2c, axis=1) / c
    spatial = False
  elif x1.ndim == 4:
    y = x1 if x2c is None else x2c
    # tensordot over channels, then zip (h,h'),(w,w')    requirements.py:548-552
    nngp = np.einsum('ahwc,bxyc->abhxwy', x1, y, optimize=True) / c
    # batched dot_general + zip                           requirements.py:534-538
    cov1 = np.einsum('ahwc,axyc->ahxwy', x1, x1, optimize=True) / c
    cov2 = None if x2c is None else np.einsum('ahwc,axyc->ahxwy', x2c, x2c, optimize=True) / c
    spatial = True
  else:
    raise NotImplementedError('oracle covers [N,d] and NHWC inputs only')
  ntk = np.zeros((), dtype) if compute_ntk else None          # requirements.py:807
  return OState(nngp=nngp, ntk=ntk, cov1=cov1, cov2=cov2, is_gaussian=False,
                is_reversed=False, spatial=spatial, shape1=tuple(x1.shape),
                shape2=tuple(x1.shape if x2 is None else x2.shape))


# --------------------------------------------------------------------------
# Dense: linear.py:899-926 with _affine linear.py:2938-2971 (ntk param only).
# --------------------------------------------------------------------------
def _affine(mat, W_std, b_std):
  if mat is None:
    return None
  mat = mat * W_std**2
  if b_std is not None:
    mat = mat + b_std**2
  return mat


def dense(st: OState, W_std: float, b_std: Optional[float]) -> OState:
  cov1, nngp, cov2 = (_affine(m, W_std, b_std) for m in (st.cov1, st.nngp, st.cov2))
  ntk = st.ntk
  if ntk is not None:
    ntk = nngp + W_std**2 * ntk                                # linear.py:911-912
  return st.replace(cov1=cov1, nngp=nngp, cov2=cov2, ntk=ntk, is_gaussian=True)


# --------------------------------------------------------------------------
# Conv: linear.py:1321-1424 -> _conv_kernel_full_spatial_shared 3115-3207 ->
# _conv_kernel_full_spatial_loop 3341-3378.  Per spatial pair (d,d') the
# reference runs a depthwise conv with rhs = diag(1/k) (3352-3363), i.e.
#   out[a,a'] = (1/k) sum_j in~[s a + j - lo, s a' + j - lo].
# --------------------------------------------------------------------------
def _diag_conv_pair(m: np.ndarray, ax: int, k: int, s: int, padding: str) -> np.ndarray:
  """Applies the diagonal-offset box filter on axes (ax, ax+1)."""
  n = m.shape[ax]
  assert m.shape[ax + 1] == n
  if padding == 'CIRCULAR':
    # linear.py:3158-3165, 2974-3019: wrap-pad by the SAME amounts, then VALID.
    _, lo, hi = same_pads(n, k, s)
    pads = [(0, 0)] * m.ndim
    pads[ax] = pads[ax + 1] = (lo, hi)
    m = np.pad(m, pads, mode='wrap')
    out = valid_out(n + lo + hi, k, s)
  elif padding == 'SAME':
    out, lo, hi = same_pads(n, k, s)
    pads = [(0, 0)] * m.ndim
    pads[ax] = pads[ax + 1] = (lo, hi)
    m = np.pad(m, pads, mode='constant')
  elif padding == 'VALID':
    out = valid_out(n, k, s)
  else:
    raise ValueError(padding)
  acc = None
  for j in range(k):
    sl = [slice(None)] * m.ndim
    sl[ax] = sl[ax + 1] = slice(j, j + (out - 1) * s + 1, s)
    term = m[tuple(sl)]
    acc = term.copy() if acc is None else acc + term
  return acc / k


def _conv_unscaled(m, filter_shape, strides, padding, batch_ndim):
  if m is None or m.ndim == 0:                                 # linear.py:3155
    return m
  # Loop order as the reference: last pair first (linear.py:3365).
  m = _diag_conv_pair(m, batch_ndim + 2, filter_shape[1], strides[1], padding)
  m = _diag_conv_pair(m, batch_ndim, filter_shape[0], strides[0], padding)
  return m


def conv(st: OState, filter_shape, strides, padding, W_std, b_std) -> OState:
  if not st.spatial:
    raise ValueError('Conv needs spatial inputs')
  b2 = None if b_std is None else b_std**2

  def cv(m, batch_ndim):                                       # linear.py:1387-1391
    out = _conv_unscaled(m, filter_shape, strides, padding, batch_ndim)
    if out is not None:
      out = out * W_std**2
      if b2 is not None:
        out = out + b2
    return out

  cov1 = cv(st.cov1, 1)
  cov2 = cv(st.cov2, 1)
  nngp = cv(st.nngp, 2)
  ntk = st.ntk
  if ntk is not None:                                          # linear.py:1396-1398
    ntk = W_std**2 * _conv_unscaled(ntk, filter_shape, strides, padding, 2) + nngp
  h, w = nngp.shape[2], nngp.shape[4]
  return st.replace(cov1=cov1, cov2=cov2, nngp=nngp, ntk=ntk, is_gaussian=True,
                    is_reversed=not st.is_reversed,            # linear.py:1344-1349
                    shape1=(st.shape1[0], h, w, st.shape1[-1]),
                    shape2=(st.shape2[0], h, w, st.shape2[-1]))


# --------------------------------------------------------------------------
# Activations.  get_diagonal_outer_prods: requirements.py:1077-1117.
# --------------------------------------------------------------------------
def _diag(cov: np.ndarray) -> np.ndarray:
  """q[n,h,w] = cov[n,h,h,w,w]  (requirements.py:1057-1074); identity for [n]."""
  if cov.ndim == 1:
    return cov
  return np.einsum('nhhww->nhw', cov)


def _outer(qa: np.ndarray, qb: np.ndarray, batch_outer: bool) -> np.ndarray:
  """prod[n1,n2,h,h',w,w'] = qa[n1,h,w]*qb[n2,h',w'] (utils.py:448-453)."""
  if qa.ndim == 1:
    return qa[:, None] * qb[None, :] if batch_outer else qa * qb
  if batch_outer:
    return np.einsum('ahw,bxy->abhxwy', qa, qb)
  return np.einsum('ahw,axy->ahxwy', qa, qb)


def _sqrt(x):                                                  # elementwise.py:1278-1280
  return np.sqrt(np.maximum(x, 0.))


def _arctan2(x, y, fill_zero):                                 # elementwise.py:1293-1299
  return np.where((x == 0.) & (y == 0.), fill_zero, np.arctan2(x, y))


def abrelu(st: OState, a: float, b: float, do_stabilize: bool = False) -> OState:
  """elementwise.py:423-477."""
  if not st.is_gaussian:                                       # elementwise.py:1267-1270
    raise ValueError('The input to the activation function must be Gaussian')
  cov1, nngp, cov2, ntk = st.cov1, st.nngp, st.cov2, st.ntk
  if do_stabilize:                                             # elementwise.py:430-436
    factor = max(float(np.max(np.abs(nngp))), 1e-12)
    nngp = nngp / factor
    cov1 = cov1 / factor
    cov2 = None if cov2 is None else cov2 / factor
  q1 = _diag(cov1)
  q2 = q1 if cov2 is None else _diag(cov2)
  prod12 = _outer(q1, q2, True)
  prod11 = _outer(q1, q1, False)
  prod22 = None if cov2 is None else _outer(q2, q2, False)

  def f(k, prod, t=None):                                      # elementwise.py:444-455
    square_root = _sqrt(prod - k**2)
    angles = _arctan2(square_root, k, math.pi / 2)
    factor_ = (a - b)**2 / (2 * math.pi)
    dot_sigma = (a**2 + b**2) / 2 - factor_ * angles
    k = factor_ * square_root + dot_sigma * k
    if t is not None:
      t = t * dot_sigma
    return k, t

  nngp, ntk = f(nngp, prod12, ntk)
  cov1, _ = f(cov1, prod11)
  if cov2 is not None:
    cov2, _ = f(cov2, prod22)
  if do_stabilize:                                             # elementwise.py:471-475
    nngp = nngp * factor
    cov1 = cov1 * factor
    cov2 = None if cov2 is None else cov2 * factor
  return st.replace(cov1=cov1, nngp=nngp, cov2=cov2, ntk=ntk, is_gaussian=False)


def erf(st: OState, a: float, b: float, c: float) -> OState:
  """elementwise.py:67-112 with Kernel.__mul__/__add__ (utils/kernel.py:426-441)."""
  if not st.is_gaussian:
    raise ValueError('The input to the activation function must be Gaussian')
  var_b = b**2                                                 # k *= b  -> b**2 * mats
  cov1 = var_b * st.cov1
  nngp = var_b * st.nngp
  cov2 = None if st.cov2 is None else var_b * st.cov2
  ntk = None if st.ntk is None else var_b * st.ntk
  d1 = 1 + 2 * _diag(cov1)
  d2 = d1 if cov2 is None else 1 + 2 * _diag(cov2)
  prod12 = _outer(d1, d2, True)
  prod11 = _outer(d1, d1, False)
  prod22 = None if cov2 is None else _outer(d2, d2, False)
  factor = 2 / math.pi

  def f(k, prod, t=None):                                      # elementwise.py:84-93
    square_root = _sqrt(prod - 4 * k**2)
    k = factor * np.arctan2(2 * k, square_root)
    if t is not None:
      t = t * (2 * factor / square_root)
    return k, t

  nngp, ntk = f(nngp, prod12, ntk)
  cov1, _ = f(cov1, prod11)
  if cov2 is not None:
    cov2, _ = f(cov2, prod22)
  va, vc = a**2, c**2                                          # a * k + c
  cov1 = vc + va * cov1
  nngp = vc + va * nngp
  cov2 = None if cov2 is None else vc + va * cov2
  ntk = None if ntk is None else va * ntk
  return st.replace(cov1=cov1, nngp=nngp, cov2=cov2, ntk=ntk, is_gaussian=False)


def _elementwise_pairs(st: OState, prep1, f):
  """Shared driver of the closed-form activations below: `prep1(q)` maps the per-sample diagonals to the
  quantity whose outer combination `f` needs (get_diagonal_outer_prods, requirements.py:1077-1117), and
  `f(k, outer12_args..., t)` is the reference's `nngp_ntk_fn` (the full, non-diagonal_spatial branch)."""
  if not st.is_gaussian:                                       # elementwise.py:1267-1270
    raise ValueError('The input to the activation function must be Gaussian')
  q1 = _diag(st.cov1)
  q2 = q1 if st.cov2 is None else _diag(st.cov2)
  nngp, ntk = f(st.nngp, prep1(q1, q2, True), st.ntk)
  cov1, _ = f(st.cov1, prep1(q1, q1, False), None)
  cov2 = None
  if st.cov2 is not None:
    cov2, _ = f(st.cov2, prep1(q2, q2, False), None)
  return st.replace(cov1=cov1, nngp=nngp, cov2=cov2, ntk=ntk, is_gaussian=False)


def _outer_sum(qa: np.ndarray, qb: np.ndarray, batch_outer: bool) -> np.ndarray:
  """sum[n1,n2,h,h',w,w'] = qa[n1,h,w] + qb[n2,h',w'] (get_diagonal_outer_prods with op.add)."""
  if qa.ndim == 1:
    return qa[:, None] + qb[None, :] if batch_outer else qa + qb
  if batch_outer:
    return qa[:, None, :, None, :, None] + qb[None, :, None, :, None, :]
  return qa[:, :, None, :, None] + qb[:, None, :, None, :]


def gelu(st: OState) -> OState:
  """elementwise.py:195-263 (Gelu, full-spatial branch `nngp_ntk_fn`)."""
  def prep(qa, qb, batch_outer):
    return _outer(qa, qb, batch_outer), _outer(qa + 1, qb + 1, batch_outer)

  def f(k, prods, t):                                          # elementwise.py:225-240
    prod, prod_plus_1 = prods
    delta_squared = prod_plus_1 - k**2
    delta = _sqrt(delta_squared)
    angles = np.arctan2(k, delta)
    new_k = (k**2 + prod * delta_squared) / (prod_plus_1 * delta)
    new_k = new_k + k * angles
    new_k = new_k / (2 * math.pi)
    new_k = new_k + 0.25 * k
    if t is not None:
      second_term = 0.25 + angles / (2 * math.pi)
      first_term = 1 / delta_squared + (1 - prod) / prod_plus_1 + 1
      first_term = first_term * (k / delta / (2. * math.pi))
      t = t * (first_term + second_term)
    return new_k, t

  return _elementwise_pairs(st, prep, f)


def sin(st: OState, a: float, b: float, c: float) -> OState:
  """elementwise.py:266-320 (Sin(a, b, c) = a sin(b x + c)); Cos(a, b, c) == Sin(a, b, c + pi/2), :323-341."""
  half_a_square = a**2 / 2.

  def f(k, sum_, t):                                           # elementwise.py:294-300
    s1 = np.exp(b**2 * (-0.5 * sum_ + k))
    s2 = np.exp(b**2 * (-0.5 * sum_ - k)) * math.cos(2 * c)
    if t is not None:
      t = t * (half_a_square * b**2 * (s1 + s2))
    return half_a_square * (s1 - s2), t

  return _elementwise_pairs(st, _outer_sum, f)


def layernorm(st: OState, eps: float) -> OState:
  """linear.py:2476-2590 with the default `axis=-1` (channel axis only): every kernel is divided by
  sqrt((eps + q1)(eps + q2)) (`get_diagonal_outer_prods(eps + cov1, eps + cov2, ..., op.mul)`, :2566-2584)."""
  if not st.is_gaussian:                                       # linear.py:2519-2521
    raise NotImplementedError('LayerNorm only implemented for Gaussian inputs.')
  q1 = eps + _diag(st.cov1)
  q2 = q1 if st.cov2 is None else eps + _diag(st.cov2)
  nngp = st.nngp / np.sqrt(_outer(q1, q2, True))
  ntk = st.ntk if (st.ntk is None or st.ntk.ndim == 0) else st.ntk / np.sqrt(_outer(q1, q2, True))
  cov1 = st.cov1 / np.sqrt(_outer(q1, q1, False))
  cov2 = None if st.cov2 is None else st.cov2 / np.sqrt(_outer(q2, q2, False))
  return st.replace(cov1=cov1, nngp=nngp, cov2=cov2, ntk=ntk)


def rbf(st: OState, gamma: float) -> OState:
  """elementwise.py:344-400 (Rbf): nngp = exp(gamma (-(q1 + q2) + 2 k)), ntk *= 2 gamma nngp."""
  def f(k, sum_, t):                                           # elementwise.py:375-379
    k = np.exp(gamma * (-sum_ + 2 * k))
    if t is not None:
      t = t * (2 * gamma * k)
    return k, t

  return _elementwise_pairs(st, _outer_sum, f)


# --------------------------------------------------------------------------
# AvgPool: linear.py:1631-1664 -> _pool_kernel 3499-3559, _normalize 3562-3572.
# reduce_window with window (wh,wh,ww,ww): independent offsets on both members.
# --------------------------------------------------------------------------
def _window_sum_axis(m: np.ndarray, ax: int, k: int, s: int, padding: str) -> np.ndarray:
  n = m.shape[ax]
  if padding == 'CIRCULAR':
    _, lo, hi = same_pads(n, k, s)
    pads = [(0, 0)] * m.ndim
    pads[ax] = (lo, hi)
    m = np.pad(m, pads, mode='wrap')
    out = valid_out(n + lo + hi, k, s)
  elif padding == 'SAME':
    out, lo, hi = same_pads(n, k, s)
    pads = [(0, 0)] * m.ndim
    pads[ax] = (lo, hi)
    m = np.pad(m, pads, mode='constant')
  else:
    out = valid_out(n, k, s)
  acc = None
  for j in range(k):
    sl = [slice(None)] * m.ndim
    sl[ax] = slice(j, j + (out - 1) * s + 1, s)
    term = m[tuple(sl)]
    acc = term.copy() if acc is None else acc + term
  return acc


def _pool_mat(m, window, strides, padding, normalize_edges, batch_ndim, pool_sum=False):
  if m is None or m.ndim == 0:                                 # linear.py:1647
    return m
  if pool_sum:                                                 # _Pooling.SUM: linear.py:3556-3557 skips _normalize
    out = m
    for i, (k, s) in enumerate(((window[0], strides[0]), (window[0], strides[0]),
                                (window[1], strides[1]), (window[1], strides[1]))):
      out = _window_sum_axis(out, batch_ndim + i, k, s, padding)
    return out
  out = m
  for i, (k, s) in enumerate(((window[0], strides[0]), (window[0], strides[0]),
                              (window[1], strides[1]), (window[1], strides[1]))):
    out = _window_sum_axis(out, batch_ndim + i, k, s, padding)
  if padding == 'SAME' and normalize_edges:                    # linear.py:3563-3569
    ones = np.ones_like(m)
    cnt = ones
    for i, (k, s) in enumerate(((window[0], strides[0]), (window[0], strides[0]),
                                (window[1], strides[1]), (window[1], strides[1]))):
      cnt = _window_sum_axis(cnt, batch_ndim + i, k, s, padding)
    return out / cnt
  return out / float(window[0]**2 * window[1]**2)              # linear.py:3570-3571


def avgpool(st: OState, window, strides, padding, normalize_edges=False, pool_sum=False) -> OState:
  """AvgPool (linear.py:1459-1501) and, with `pool_sum`, SumPool (linear.py:1503-1547): same rule without the
  division by the window."""
  if not st.spatial:
    raise ValueError('AvgPool needs spatial inputs')
  nngp = _pool_mat(st.nngp, window, strides, padding, normalize_edges, 2, pool_sum)
  ntk = _pool_mat(st.ntk, window, strides, padding, normalize_edges, 2, pool_sum)
  cov1 = _pool_mat(st.cov1, window, strides, padding, normalize_edges, 1, pool_sum)
  cov2 = _pool_mat(st.cov2, window, strides, padding, normalize_edges, 1, pool_sum)
  h, w = nngp.shape[2], nngp.shape[4]
  return st.replace(nngp=nngp, ntk=ntk, cov1=cov1, cov2=cov2,
                    shape1=(st.shape1[0], h, w, st.shape1[-1]),
                    shape2=(st.shape2[0], h, w, st.shape2[-1]))


# --------------------------------------------------------------------------
# GlobalAvgPool: linear.py:1771-1801 (+ mean_and_var requirements.py:1120-1159)
# Flatten: linear.py:1865-1899 (trace/size loop 1880-1882)
# --------------------------------------------------------------------------
def global_avg_pool(st: OState, pool_sum: bool = False) -> OState:
  """GlobalAvgPool (linear.py:1723-1808); `pool_sum`: GlobalSumPool (linear.py:1674-1720, `jnp.sum`, :1748)."""
  if not st.spatial:
    raise ValueError('GlobalAvgPool needs spatial inputs')

  def mp(m, batch_ndim):
    if m is None:
      return m
    if m.ndim == 0:
      return m                       # mean / sum over no axes of a 0-d array is itself
    ax = tuple(range(batch_ndim, m.ndim))
    return m.sum(axis=ax) if pool_sum else m.mean(axis=ax)

  return st.replace(nngp=mp(st.nngp, 2), ntk=mp(st.ntk, 2), cov1=mp(st.cov1, 1),
                    cov2=mp(st.cov2, 1), spatial=False, is_reversed=False,
                    shape1=(st.shape1[0], st.shape1[-1]),
                    shape2=(st.shape2[0], st.shape2[-1]))


def flatten(st: OState) -> OState:
  def tr(m, batch_ndim):
    if m is None or m.ndim == 0:
      return m
    while m.ndim > batch_ndim:
      m = np.trace(m, axis1=-2, axis2=-1) / m.shape[-1]
    return m

  if st.spatial:
    n_feat = int(np.prod(st.shape1[1:]))
    out = st.replace(nngp=tr(st.nngp, 2), ntk=tr(st.ntk, 2), cov1=tr(st.cov1, 1),
                     cov2=tr(st.cov2, 1), spatial=False,
                     shape1=(st.shape1[0], n_feat), shape2=(st.shape2[0], n_feat))
  else:
    out = st
  return out.replace(is_gaussian=False, is_reversed=False)


# --------------------------------------------------------------------------
# FanOut / parallel / FanInSum: branching.py:36-117, 376-411; combinators.py.
# --------------------------------------------------------------------------
def fan_in_sum(sts: Sequence[OState]) -> OState:
  if not all(s.shape1 == sts[0].shape1 and s.shape2 == sts[0].shape2 for s in sts[1:]):
    raise ValueError('All shapes should be equal in `FanInSum`')   # branching.py:71-75
  is_gaussian = all(s.is_gaussian for s in sts)
  if not is_gaussian and len(sts) != 1:                            # branching.py:77-85
    raise NotImplementedError('`FanInSum` is only implemented for Gaussian inputs')
  n_rev = sum(s.is_reversed for s in sts)                          # branching.py:391-404
  is_reversed = n_rev > len(sts) / 2

  def sm(ms):
    return None if ms[0] is None else sum(ms[1:], ms[0])

  return sts[0].replace(nngp=sm([s.nngp for s in sts]), ntk=sm([s.ntk for s in sts]),
                        cov1=sm([s.cov1 for s in sts]), cov2=sm([s.cov2 for s in sts]),
                        is_gaussian=is_gaussian, is_reversed=is_reversed)


# --------------------------------------------------------------------------
# Spec interpreter (serial: combinators.py:60-66; parallel: 194-196).
# --------------------------------------------------------------------------
def apply_spec(spec, st):
  kind = spec[0]
  if kind == 'serial':
    for s in spec[1]:
      st = apply_spec(s, st)
    return st
  if kind == 'parallel':
    if not isinstance(st, list) or len(st) != len(spec[1]):
      raise ValueError('parallel expects a list of kernels')
    return [apply_spec(s, k) for s, k in zip(spec[1], st)]
  if kind == 'fanout':
    return [st] * spec[1]
  if kind == 'faninsum':
    return fan_in_sum(st)
  if kind == 'identity':
    return st
  if kind == 'dense':
    return dense(st, spec[1], spec[2])
  if kind == 'conv':
    return conv(st, spec[1], spec[2], spec[3], spec[4], spec[5])
  if kind == 'abrelu':
    return abrelu(st, spec[1], spec[2], spec[3] if len(spec) > 3 else False)
  if kind == 'erf':
    return erf(st, spec[1], spec[2], spec[3])
  if kind == 'avgpool':
    return avgpool(st, spec[1], spec[2], spec[3], spec[4] if len(spec) > 4 else False)
  if kind == 'gap':
    return global_avg_pool(st)
  if kind == 'sumpool':
    return avgpool(st, spec[1], spec[2], spec[3], False, True)
  if kind == 'gsp':
    return global_avg_pool(st, True)
  if kind == 'gelu':
    return gelu(st)
  if kind == 'sin':
    return sin(st, spec[1], spec[2], spec[3])
  if kind == 'cos':
    return sin(st, spec[1], spec[2], spec[3] + math.pi / 2)
  if kind == 'rbf':
    return rbf(st, spec[1])
  if kind == 'layernorm':
    return layernorm(st, spec[1])
  if kind == 'flatten':
    return flatten(st)
  raise ValueError(f'unknown spec {kind}')


def kernel_state(spec, x1, x2=None, compute_ntk=True, dtype=np.float64) -> OState:
  st = inputs_to_state(np.asarray(x1), None if x2 is None else np.asarray(x2),
                       compute_ntk, dtype)
  return apply_spec(spec, st)


def kernel_fn(spec, x1, x2=None, get=('nngp', 'ntk'), dtype=np.float64):
  """Oracle analogue of `kernel_fn(x1, x2, get)`; returns a tuple in `get` order
  (or a single array when `get` is a string)."""
  names = (get,) if isinstance(get, str) else tuple(get)
  st = kernel_state(spec, x1, x2, 'ntk' in names, dtype)
  vals = []
  for n in names:
    v = getattr(st, n)
    if n == 'ntk' and v is not None and np.ndim(v) == 0:
      v = np.broadcast_to(v, st.nngp.shape).copy()
    vals.append(v)
  return vals[0] if isinstance(get, str) else tuple(vals)


def to_reference_layout(m: Optional[np.ndarray], is_reversed: bool, batch_ndim: int):
  """Canonical [.., h,h', w,w'] -> reference storage ([.., w,w', h,h'] if reversed)
  (utils/kernel.py:169-189, utils.py:465-480)."""
  if m is None or m.ndim < batch_ndim + 4 or not is_reversed:
    return m
  b = batch_ndim
  perm = tuple(range(b)) + (b + 2, b + 3, b, b + 1)
  return np.transpose(m, perm)


# --------------------------------------------------------------------------
# nt.batch tiling arithmetic: batching.py:647-679.
# --------------------------------------------------------------------------
def n_batches_and_batch_sizes(n1: int, n2: int, batch_size: int, device_count: int):
  max_serial = math.gcd(n1, n2) // device_count
  n2_bs = min(batch_size, max_serial)
  n1_bs = n2_bs * device_count
  n1_batches, ragged = divmod(n1, n1_bs)
  if ragged:
    raise ValueError('Number of rows of kernel must divide batch size.')
  n2_batches, ragged = divmod(n2, n2_bs)
  if ragged:
    raise ValueError('Number of columns of kernel must divide batch size.')
  return n1_batches, n1_bs, n2_batches, n2_bs


# --------------------------------------------------------------------------
# Convenience network specs used by tests / bench (configs of BASELINE.json).
# --------------------------------------------------------------------------
def myrtle_spec(depth: int, W_std: float = 2**0.5, b_std: float = 0.,
                tail: str = 'flatten') -> tuple:
  """notebooks/myrtle_kernel_with_neural_tangents.ipynb:114-128."""
  factor = {5: [2, 1, 1], 7: [2, 2, 2], 10: [3, 3, 3]}[depth]
  cv = ('conv', (3, 3), (1, 1), 'SAME', W_std, b_std)
  relu = ('abrelu', 0., 1., False)
  pool = ('avgpool', (2, 2), (2, 2), 'VALID', False)
  layers: List[Any] = []
  layers += [cv, relu] * factor[0] + [pool]
  layers += [cv, relu] * factor[1] + [pool]
  layers += [cv, relu] * factor[2]
  if tail == 'flatten':
    layers += [pool] * 3 + [('flatten',)]
  else:
    layers += [('gap',)]
  layers += [('dense', W_std, b_std)]
  return ('serial', layers)


def fcn_spec(depth: int = 3, W_std: float = 2., b_std: float = 0.05) -> tuple:
  """examples/infinite_fcn.py:42-46 deepened (BASELINE configs[0])."""
  layers: List[Any] = []
  for _ in range(depth):
    layers += [('dense', W_std, b_std), ('abrelu', 0., 1., False)]
  layers += [('dense', W_std, b_std)]
  return ('serial', layers)
