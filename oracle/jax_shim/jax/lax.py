"""NumPy float64 implementations of the `jax.lax` primitives on the hot path.

Semantics follow the public XLA/JAX operation definitions:
  * padtype_to_pads: SAME -> out=ceil(n/s), pad=max((out-1)s+k-n,0), lo=pad//2.
  * conv_general_dilated: cross-correlation (no kernel flip), NCHW/OIHW/NCHW,
    grouped by `feature_group_count`.
  * reduce_window: sliding-window reduction with `init_value` padding.
  * dot_general: contraction with batch dims first, then lhs free, then rhs free.
"""
import enum
import itertools
import numpy as _np

from .numpy import _wrap


class Precision(enum.Enum):
  DEFAULT = 0
  HIGH = 1
  HIGHEST = 2


class DotDimensionNumbers(tuple):
  pass


@_wrap
def add(x, y):
  return x + y


def padtype_to_pads(in_shape, window_shape, window_strides, padding):
  padding = padding.upper() if isinstance(padding, str) else padding.name
  if padding == 'SAME':
    pads = []
    for n, k, s in zip(in_shape, window_shape, window_strides):
      out = -(-n // s)
      tot = max((out - 1) * s + k - n, 0)
      pads.append((tot // 2, tot - tot // 2))
    return pads
  if padding == 'VALID':
    return [(0, 0)] * len(in_shape)
  raise ValueError(padding)


@_wrap
def reduce_window(operand, init_value, computation, window_dimensions,
                  window_strides, padding):
  assert computation is add, 'only additive windows are used by the reference'
  x = _np.asarray(operand)
  if isinstance(padding, str):
    pads = padtype_to_pads(x.shape, window_dimensions, window_strides, padding)
  else:
    pads = list(padding)
  x = _np.pad(x, pads, mode='constant', constant_values=float(_np.asarray(init_value)))
  out_shape = tuple((n - k) // s + 1 for n, k, s in
                    zip(x.shape, window_dimensions, window_strides))
  out = _np.zeros(out_shape, x.dtype)
  for offs in itertools.product(*[range(k) for k in window_dimensions]):
    sl = tuple(slice(o, o + (m - 1) * s + 1, s)
               for o, m, s in zip(offs, out_shape, window_strides))
    out = out + x[sl]
  return out


def reduce_window_shape_tuple(operand_shape, window_dimensions, window_strides, padding,
                              base_dilation=None, window_dilation=None):
  return tuple((n + lo + hi - k) // s + 1 for n, k, s, (lo, hi) in
               zip(operand_shape, window_dimensions, window_strides, padding))


def _conv_nchw(lhs, rhs, window_strides, padding, feature_group_count):
  n, c, h, w = lhs.shape
  o, i, kh, kw = rhs.shape
  g = feature_group_count
  assert c == i * g and o % g == 0
  if isinstance(padding, str):
    pads = padtype_to_pads((h, w), (kh, kw), window_strides, padding)
  else:
    pads = list(padding)
  x = _np.pad(lhs, [(0, 0), (0, 0)] + list(pads), mode='constant')
  sh, sw = window_strides
  oh = (x.shape[2] - kh) // sh + 1
  ow = (x.shape[3] - kw) // sw + 1
  out = _np.zeros((n, o, oh, ow), _np.result_type(lhs.dtype, rhs.dtype))
  opg = o // g
  for grp in range(g):
    xg = x[:, grp * i:(grp + 1) * i]                       # [n, i, H, W]
    wg = rhs[grp * opg:(grp + 1) * opg]                    # [opg, i, kh, kw]
    for a in range(kh):
      for b in range(kw):
        wt = wg[:, :, a, b]                                # [opg, i]
        if not wt.any():
          continue
        patch = xg[:, :, a:a + (oh - 1) * sh + 1:sh, b:b + (ow - 1) * sw + 1:sw]
        out[:, grp * opg:(grp + 1) * opg] += _np.einsum('oi,nihw->nohw', wt, patch)
  return out


@_wrap
def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None,
                         rhs_dilation=None, dimension_numbers=None,
                         feature_group_count=1, batch_group_count=1,
                         precision=None, preferred_element_type=None):
  assert lhs_dilation is None or all(d == 1 for d in lhs_dilation)
  assert rhs_dilation is None or all(d == 1 for d in rhs_dilation)
  assert batch_group_count == 1
  lhs = _np.asarray(lhs)
  rhs = _np.asarray(rhs)
  if dimension_numbers is None:
    dimension_numbers = ('NCHW', 'OIHW', 'NCHW')
  lhs_spec, rhs_spec, out_spec = dimension_numbers
  assert lhs.ndim == 4, 'shim supports 2 spatial dimensions'
  sp = [ch for ch in rhs_spec if ch not in 'OI']           # spatial letters, filter order
  lperm = [lhs_spec.index('N'), lhs_spec.index('C')] + [lhs_spec.index(ch) for ch in sp]
  rperm = [rhs_spec.index('O'), rhs_spec.index('I')] + [rhs_spec.index(ch) for ch in sp]
  out = _conv_nchw(lhs.transpose(lperm), rhs.transpose(rperm), tuple(window_strides),
                   padding, feature_group_count)
  canon = 'NC' + ''.join(sp)
  return out.transpose([canon.index(ch) for ch in out_spec])


def conv_transpose(*a, **k):
  raise NotImplementedError


def conv_general_dilated_local(*a, **k):
  raise NotImplementedError


@_wrap
def dot_general(lhs, rhs, dimension_numbers, precision=None, preferred_element_type=None):
  (lc, rc), (lb, rb) = dimension_numbers
  lhs = _np.asarray(lhs)
  rhs = _np.asarray(rhs)
  lc, rc, lb, rb = map(lambda t: tuple(int(v) % 1000000 for v in t), (lc, rc, lb, rb))
  letters = iter('abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ')
  l_sub = [None] * lhs.ndim
  r_sub = [None] * rhs.ndim
  batch = []
  for a, b in zip(lb, rb):
    ch = next(letters)
    l_sub[a] = r_sub[b] = ch
    batch.append(ch)
  for a, b in zip(lc, rc):
    ch = next(letters)
    l_sub[a] = r_sub[b] = ch
  l_free, r_free = [], []
  for ax in range(lhs.ndim):
    if l_sub[ax] is None:
      l_sub[ax] = next(letters)
      l_free.append(l_sub[ax])
  for ax in range(rhs.ndim):
    if r_sub[ax] is None:
      r_sub[ax] = next(letters)
      r_free.append(r_sub[ax])
  expr = f"{''.join(l_sub)},{''.join(r_sub)}->{''.join(batch + l_free + r_free)}"
  return _np.einsum(expr, lhs, rhs, optimize=True)


def scan(f, init, xs, length=None):
  raise NotImplementedError


def cond(pred, true_fun, false_fun, *operands):
  return true_fun(*operands) if pred else false_fun(*operands)
