"""`stax`-compatible front end: layers return `(init_fn, apply_fn, kernel_fn)`.

Mirrors the public surface of `neural_tangents/stax.py:77-165` for the in-scope
layers.  `kernel_fn(x1_or_kernel, x2=None, get=None, *, pattern=None,
mask_constant=None, diagonal_batch=None, diagonal_spatial=None)` has the
reference's signature (`_src/stax/requirements.py:955-965`); the layer tree is
lowered to a slot program and executed by the CUDA library
(`include/ntk_b200.h`).  Anything the reference supports but this path does not
raises `NotImplementedError` instead of silently mis-computing.

`init_fn` / `apply_fn` are NumPy finite-width implementations, used for shape
inference (`shape1`/`shape2`, `requirements.py:833-879`) and Monte-Carlo checks.
"""
import collections
import math
import threading
from typing import Optional

import numpy as np

from . import _lib
from ._config import config
from .kernel import Kernel

import enum

__all__ = ['serial', 'parallel', 'FanOut', 'FanInSum', 'Identity', 'Dense', 'Conv', 'Relu',
           'ABRelu', 'LeakyRelu', 'Abs', 'Erf', 'Sigmoid_like', 'Gelu', 'Sin', 'Cos', 'Rbf', 'AvgPool', 'SumPool',
           'GlobalAvgPool', 'GlobalSumPool', 'Flatten', 'LayerNorm', 'Padding', 'Bool', 'Diagonal']


class Padding(enum.Enum):
  """`_src/stax/linear.py:54-69`."""
  CIRCULAR = 'CIRCULAR'
  SAME = 'SAME'
  VALID = 'VALID'


# Names the reference's `stax` exports that are outside the B200 hot path: asking
# for them fails loudly instead of with an AttributeError that looks like a typo.
_OUT_OF_SCOPE = ('repeat', 'Elementwise', 'ElementwiseNumerical', 'Exp', 'ExpNormalized',
                 'Gabor', 'Gaussian', 'Hermite', 'Monomial', 'Polynomial',
                 'RectifiedMonomial', 'Sign', 'Aggregate', 'ConvLocal',
                 'ConvTranspose', 'Index', 'DotGeneral', 'Dropout', 'GlobalSelfAttention',
                 'ImageResize', 'Slice', 'FanInConcat',
                 'FanInProd', 'AggregateImplementation', 'AttentionMechanism', 'PositionalEmbedding',
                 'MaskedArray', 'layer', 'requires', 'supports_masking', 'unmask_fn')


def __getattr__(name):
  if name in _OUT_OF_SCOPE:
    raise NotImplementedError(f'stax.{name} exists in neural_tangents but is outside the B200 '
                              'hot path (Dense/Conv/Relu/Erf/Gelu/Sin/Cos/Rbf/AvgPool/SumPool/GlobalAvgPool/'
                              'GlobalSumPool/Flatten/FanOut/FanInSum/serial/parallel); see DESIGN.md §9')
  raise AttributeError(name)


# ----------------------------------------------------------------------------
# frozen mapping for `kernel_fn.input_req` (the reference uses frozendict)
# ----------------------------------------------------------------------------
class FrozenDict(dict):
  def __hash__(self):
    return hash(frozenset(self.items()))

  def _ro(self, *a, **k):
    raise TypeError('immutable')

  __setitem__ = __delitem__ = clear = pop = popitem = setdefault = update = _ro


_DEFAULT_INPUT_REQ = FrozenDict(diagonal_batch=True, diagonal_spatial=False, batch_axis=0,
                                use_dropout=False, channel_axis=-1, mask_constant=None)


class Bool(enum.IntEnum):
  """Trinary logic of the requirement negotiation (`_src/stax/requirements.py:397-422`)."""
  NO = 0
  MAYBE = 1
  YES = 2

  def __and__(self, other):
    return Bool(min(int(self), int(other)))

  __rand__ = __and__


class Diagonal:
  """Whether a layer can work on / emits kernels that hold only the spatial diagonal
  (`_src/stax/requirements.py:425-515`).  `a >> b` is the requirement of `b(a(.))` (serial), `a & b` that of two
  parallel branches; `bool(d)` is the `diagonal_spatial` the network runs with when the user does not ask."""
  __slots__ = ('input', 'output')

  def __init__(self, input: Bool = Bool.YES, output: Bool = Bool.NO):
    object.__setattr__(self, 'input', Bool(input))
    object.__setattr__(self, 'output', Bool(output))

  def __setattr__(self, *a):
    raise AttributeError('Diagonal is immutable')

  def __rshift__(self, other: 'Diagonal') -> 'Diagonal':
    if self.output == Bool.YES:          # everything behind is diagonal already: later layers cannot constrain
      return self
    if self.output > Bool.NO and other.input > Bool.NO:
      inp = self.input
    elif self.output == Bool.NO and other.input < Bool.YES:
      inp = Bool.NO
    else:
      inp = Bool(min(self.input, other.input))
    return Diagonal(inp, other.output)

  def __lshift__(self, other: 'Diagonal') -> 'Diagonal':
    return other >> self

  def __and__(self, other: 'Diagonal') -> 'Diagonal':
    return Diagonal(self.input & other.input, self.output & other.output)

  __rand__ = __and__

  def __bool__(self):
    return self.input == Bool.YES and self.output > Bool.NO

  def __eq__(self, other):
    return isinstance(other, Diagonal) and (self.input, self.output) == (other.input, other.output)

  def __hash__(self):
    return hash((int(self.input), int(self.output)))

  def __repr__(self):
    return f'Diagonal(input={self.input!r}, output={self.output!r})'


def _fold_input_req(kernel_fns, fold):
  """`_get_input_req_attr` (`_src/stax/combinators.py:204-268`): requirements of a `serial` (fold = `>>`) or
  `parallel` (fold = `&`) combination from those of its members."""
  import operator
  import warnings
  req = {}
  for f in kernel_fns:
    for k, v in getattr(f, 'input_req', {}).items():
      if k == 'use_dropout':
        if k in req and req[k] != v:
          raise ValueError('`use_dropout` is a single whole-network attribute and cannot be set to different values.')
        req[k] = v
      elif k in ('batch_axis', 'channel_axis'):
        if k not in req:
          req[k] = v
        elif fold is operator.and_ and req[k] != v:
          warnings.warn(f'For `kernel_fn`, `{k}` parameters must match in all parallel branches, got {req[k]} and {v}.')
      elif k in ('diagonal_batch', 'diagonal_spatial'):
        req[k] = fold(req[k], v) if k in req else v
      else:
        raise NotImplementedError(k)
  return req

_KERNEL_FIELDS = ('nngp', 'ntk', 'cov1', 'cov2', 'x1_is_x2', 'is_gaussian', 'is_reversed',
                  'is_input', 'diagonal_batch', 'diagonal_spatial', 'shape1', 'shape2',
                  'batch_axis', 'channel_axis', 'mask1', 'mask2')
_NT_CACHE = {}


def _canonicalize_get(get):
  """`_src/utils/utils.py:139-155`."""
  if get is None:
    return True, None
  if not get:
    raise ValueError('"get" must be non-empty.')
  is_str = isinstance(get, str)
  if is_str:
    get = (get,)
  get = tuple(s.lower() for s in get)
  if len(set(get)) < len(get):
    raise ValueError('All entries in "get" must be unique. Got {}'.format(get))
  for g in get:
    if g not in _KERNEL_FIELDS:
      raise ValueError(f'unknown kernel field {g!r} in "get"')
  return is_str, get


def _namedtuple(get):
  if get not in _NT_CACHE:
    _NT_CACHE[get] = collections.namedtuple('AnalyticKernel', get)
  return _NT_CACHE[get]


def _x1_is_x2(x1, x2, eps=1e-12):
  """`_src/utils/utils.py:255-282`."""
  if x2 is None or x1 is x2:
    return True
  if x1.shape != x2.shape:
    return False
  return bool(np.all(np.abs(x1 - x2) < eps))


# ----------------------------------------------------------------------------
# geometry helpers (lax.padtype_to_pads semantics)
# ----------------------------------------------------------------------------
def _axis_out(n, k, s, padding):
  if padding == 'VALID':
    return (n - k) // s + 1 if n >= k else 0
  return -(-n // s)


def _same_pads(n, k, s):
  out = -(-n // s)
  tot = max((out - 1) * s + k - n, 0)
  return tot // 2, tot - tot // 2


# ----------------------------------------------------------------------------
# finite-width NumPy layers (shape inference + Monte-Carlo checks)
# ----------------------------------------------------------------------------
def _rng(rng):
  if isinstance(rng, np.random.Generator):
    return rng
  if rng is None:
    return np.random.default_rng(0)
  return np.random.default_rng(np.asarray(rng).astype(np.uint32).ravel())


def _pad_spatial(x, pads, padding):
  cfg = [(0, 0), pads[0], pads[1], (0, 0)]
  return np.pad(x, cfg, mode='wrap' if padding == 'CIRCULAR' else 'constant')


def _windows(x, k, s, padding):
  """NHWC -> [N, Ho, Wo, kh, kw, C] strided view of (padded) x."""
  if padding in ('SAME', 'CIRCULAR'):
    x = _pad_spatial(x, (_same_pads(x.shape[1], k[0], s[0]), _same_pads(x.shape[2], k[1], s[1])),
                     padding)
  n, h, w, c = x.shape
  ho, wo = (h - k[0]) // s[0] + 1, (w - k[1]) // s[1] + 1
  st = x.strides
  return np.lib.stride_tricks.as_strided(
      x, (n, ho, wo, k[0], k[1], c), (st[0], st[1] * s[0], st[2] * s[1], st[1], st[2], st[3]),
      writeable=False)


# ----------------------------------------------------------------------------
# spec -> slot program lowering
# ----------------------------------------------------------------------------
class _Meta:
  __slots__ = ('is_reversed', 'ntk_tensor', 'spatial')

  def __init__(self, is_reversed=False, ntk_tensor=False, spatial=True):
    self.is_reversed, self.ntk_tensor, self.spatial = is_reversed, ntk_tensor, spatial

  def copy(self):
    return _Meta(self.is_reversed, self.ntk_tensor, self.spatial)


class _Lowered:
  def __init__(self, spec, in_reversed, in_ntk_tensor, in_spatial):
    self.ops = []
    self.meta = [_Meta(in_reversed, in_ntk_tensor, in_spatial)]
    out = self._lower(spec, 0)
    if isinstance(out, list):
      raise ValueError('network output is a list of kernels (FanOut without FanIn); '
                       'NTTree outputs are not supported on this path')
    self.out_slot = out
    self.out_meta = self.meta[out]
    self.program = _lib.Program(self.ops, len(self.meta), out)

  def _new(self, meta):
    self.meta.append(meta)
    return len(self.meta) - 1

  def _emit(self, kind, src, ints=(), floats=(), src2=-1, meta=None):
    dst = self._new(meta)
    self.ops.append((kind, src, src2, dst, tuple(ints), tuple(floats)))
    return dst

  def _lower(self, spec, cur):
    kind = spec[0]
    if kind == 'serial':
      for s in spec[1]:
        cur = self._lower(s, cur)
      return cur
    if kind == 'fanout':
      if isinstance(cur, list):
        raise NotImplementedError('nested FanOut on a list of kernels')
      return [cur] * spec[1]
    if kind == 'parallel':
      if not isinstance(cur, list) or len(cur) != len(spec[1]):
        raise ValueError('`parallel` expects one input kernel per branch (use FanOut first)')
      return [self._lower(s, c) for s, c in zip(spec[1], cur)]
    if kind == 'faninsum':
      if not isinstance(cur, list):
        cur = [cur]
      metas = [self.meta[c] for c in cur]
      # branching.py:391-404: majority vote on is_reversed
      is_rev = sum(m.is_reversed for m in metas) > len(metas) / 2
      acc = cur[0]
      for c in cur[1:]:
        m = _Meta(is_rev, self.meta[acc].ntk_tensor or self.meta[c].ntk_tensor, self.meta[acc].spatial)
        acc = self._emit(_lib.OP_FANINSUM, acc, src2=c, meta=m)
      if len(cur) == 1:
        acc = self._emit(_lib.OP_IDENTITY, acc, meta=self.meta[acc].copy())
      self.meta[acc].is_reversed = is_rev
      return acc
    if isinstance(cur, list):
      raise ValueError(f'layer {kind!r} applied to a list of kernels; wrap it in `parallel`')
    m = self.meta[cur].copy()
    if kind == 'identity':
      return cur
    if kind == 'dense':
      m.ntk_tensor = True
      return self._emit(_lib.OP_DENSE, cur, (spec[2] is not None,),
                        (spec[1]**2, 0. if spec[2] is None else spec[2]**2), meta=m)
    if kind == 'conv':
      m.ntk_tensor = True
      m.is_reversed = not m.is_reversed                      # linear.py:1344-1349
      (kh, kw), (sh, sw) = spec[1], spec[2]
      return self._emit(_lib.OP_CONV, cur, (kh, kw, sh, sw, _lib.PAD[spec[3]], spec[5] is not None),
                        (spec[4]**2, 0. if spec[5] is None else spec[5]**2), meta=m)
    if kind == 'abrelu':
      return self._emit(_lib.OP_ABRELU, cur, (bool(spec[3]),), (spec[1], spec[2]), meta=m)
    if kind == 'erf':
      return self._emit(_lib.OP_ERF, cur, (), (spec[1], spec[2], spec[3]), meta=m)
    if kind == 'gelu':
      return self._emit(_lib.OP_GELU, cur, (), (), meta=m)
    if kind == 'sin':
      return self._emit(_lib.OP_SIN, cur, (), (spec[1], spec[2], spec[3]), meta=m)
    if kind == 'rbf':
      return self._emit(_lib.OP_RBF, cur, (), (spec[1],), meta=m)
    if kind == 'layernorm':
      return self._emit(_lib.OP_LAYERNORM, cur, (), (spec[1],), meta=m)
    if kind in ('avgpool', 'sumpool'):
      (wh, ww), (sh, sw) = spec[1], spec[2]
      flags = (1 if spec[4] else 0) | (2 if kind == 'sumpool' else 0)   # bit 0 normalize_edges, bit 1 SumPool
      return self._emit(_lib.OP_AVGPOOL, cur, (wh, ww, sh, sw, _lib.PAD[spec[3]], flags), meta=m)
    if kind in ('gap', 'gsp'):
      m.is_reversed, m.spatial = False, False                # linear.py:1801
      return self._emit(_lib.OP_GAP, cur, (1 if kind == 'gsp' else 0,), meta=m)
    if kind == 'flatten':
      m.is_reversed, m.spatial = False, False                # linear.py:1896
      return self._emit(_lib.OP_FLATTEN, cur, meta=m)
    raise ValueError(f'unknown layer spec {kind!r}')


_lower_cache = {}
_lower_lock = threading.Lock()


def _freeze(spec):
  if isinstance(spec, (list, tuple)):
    return tuple(_freeze(s) for s in spec)
  return spec


def _lowered(spec, in_reversed=False, in_ntk_tensor=False, in_spatial=True):
  key = (_freeze(spec), in_reversed, in_ntk_tensor, in_spatial)
  with _lower_lock:
    low = _lower_cache.get(key)
    if low is None:
      low = _Lowered(spec, in_reversed, in_ntk_tensor, in_spatial)
      _lower_cache[key] = low
    return low


# ----------------------------------------------------------------------------
# output-shape inference of the finite network (replaces eval_shape)
# ----------------------------------------------------------------------------
def _out_shape(spec, shape):
  kind = spec[0]
  if kind == 'serial':
    for s in spec[1]:
      shape = _out_shape(s, shape)
    return shape
  if kind == 'fanout':
    return [shape] * spec[1]
  if kind == 'parallel':
    return [_out_shape(s, sh) for s, sh in zip(spec[1], shape)]
  if kind == 'faninsum':
    return shape[0] if isinstance(shape, list) else shape
  if kind in ('identity', 'abrelu', 'erf', 'gelu', 'sin', 'rbf', 'layernorm'):
    return shape
  shape = tuple(shape)
  if kind == 'dense':
    return shape[:-1] + (spec[-1],)
  if kind == 'conv':
    return (shape[0], _axis_out(shape[1], spec[1][0], spec[2][0], spec[3]),
            _axis_out(shape[2], spec[1][1], spec[2][1], spec[3]), spec[-1])
  if kind in ('avgpool', 'sumpool'):
    return (shape[0], _axis_out(shape[1], spec[1][0], spec[2][0], spec[3]),
            _axis_out(shape[2], spec[1][1], spec[2][1], spec[3]), shape[3])
  if kind in ('gap', 'gsp'):
    return (shape[0], shape[-1])
  if kind == 'flatten':
    return (shape[0], int(np.prod(shape[1:])))
  raise ValueError(kind)


# ----------------------------------------------------------------------------
# layout conversion canonical <-> reference (is_reversed)
# ----------------------------------------------------------------------------
def _to_ref_layout(m, is_reversed, batch_ndim):
  if m is None or not is_reversed or m.ndim < batch_ndim + 4:
    return m
  b = batch_ndim
  return np.ascontiguousarray(np.transpose(m, tuple(range(b)) + (b + 2, b + 3, b, b + 1)))


_from_ref_layout = _to_ref_layout  # the (H,H')<->(W,W') swap is an involution


# ----------------------------------------------------------------------------
# kernel_fn factory
# ----------------------------------------------------------------------------
def _make_kernel_fn(spec, req):
  # `spec` carries out_chan as a trailing element for dense/conv; strip for hashing/lowering
  def kernel_fn(x1_or_kernel, x2=None, get=None, *, pattern=None, mask_constant=None,
                diagonal_batch=None, diagonal_spatial=None, **kwargs):
    is_str, get_c = _canonicalize_get(get)
    if pattern is not None:
      raise NotImplementedError('`pattern` (Aggregate layers) is outside the B200 hot path')
    if mask_constant is not None:
      raise NotImplementedError('masking (`mask_constant`) is outside the B200 hot path')
    if diagonal_batch is False:
      raise NotImplementedError('`diagonal_batch=False` is outside the B200 hot path')
    if kwargs:
      raise NotImplementedError(f'unsupported kernel_fn arguments: {sorted(kwargs)}')
    d_req = req.get('diagonal_spatial')
    if diagonal_spatial is True and (d_req is False or (isinstance(d_req, Diagonal) and d_req.input == Bool.NO)):
      raise ValueError(f'Asked to compute `kernel_fn` output with `diagonal_spatial == True`, while `kernel_fn` '
                       f'requires `diagonal_spatial == {d_req}`.')

    _warn_fan_in(spec)
    if isinstance(x1_or_kernel, Kernel) and x2 is None:
      out = _apply_to_kernel(spec, x1_or_kernel)
    else:
      out = _apply_to_inputs(spec, x1_or_kernel, x2, get_c)
    if diagonal_spatial:
      out = _take_spatial_diagonal(out)
    if get_c is None:
      return out
    d = out.asdict()
    if is_str:
      return d[get_c[0]]
    return _namedtuple(get_c)(*(d[g] for g in get_c))

  kernel_fn.input_req = req
  kernel_fn._spec = spec
  return kernel_fn


def _take_spatial_diagonal(k: Kernel) -> Kernel:
  """User-requested `diagonal_spatial=True` on a spatial output: the full
  representation is always sufficient (tests/stax/requirements_test.py:87-111)."""
  if k.nngp.ndim != 6:
    return k
  d2 = lambda m: None if m is None or m.ndim != 6 else np.einsum('abhhww->abhw', m)
  d1 = lambda m: None if m is None else np.einsum('ahhww->ahw', m)
  return k.replace(nngp=d2(k.nngp), ntk=k.ntk if (k.ntk is None or k.ntk.ndim == 0) else d2(k.ntk),
                   cov1=d1(k.cov1), cov2=d1(k.cov2), diagonal_spatial=True)


def _strip(spec):
  """Removes front-end-only fields (out_chan) so equal kernels share a program."""
  kind = spec[0]
  if kind in ('serial', 'parallel'):
    return (kind, tuple(_strip(s) for s in spec[1]))
  if kind == 'dense':
    return spec[:3]
  if kind == 'conv':
    return spec[:6]
  return tuple(spec)


def _check_inputs(x1, x2):
  if not isinstance(x1, np.ndarray) or not (x2 is None or isinstance(x2, np.ndarray)):
    raise TypeError(f'Wrong input types given. Found `x1` of type {type(x1)} and `x2` of type '
                    f'{type(x2)}, need both to be `np.ndarray`s (`x2` can be `None`).')
  if x1.ndim < 2:
    raise ValueError(f'Inputs must be at least 2D (a batch dimension and a channel/feature '
                     f'dimension), got {x1.ndim}.')
  if x1.ndim not in (2, 4):
    raise NotImplementedError('the B200 hot path covers [N, d] and NHWC [N, H, W, C] inputs; '
                              f'got ndim={x1.ndim}')
  if x2 is not None and x2.shape[1:] != x1.shape[1:]:
    raise ValueError(f'x1 and x2 must agree on all non-batch dimensions, got {x1.shape} and {x2.shape}')


def _apply_to_inputs(spec, x1, x2, get_c):
  _check_inputs(x1, x2)
  dt = config.dtype
  x1c = np.ascontiguousarray(x1, dt)
  x2c = None if x2 is None else np.ascontiguousarray(x2, dt)
  spatial = x1.ndim == 4
  H, W = (x1.shape[1], x1.shape[2]) if spatial else (0, 0)
  C = x1.shape[-1]
  low = _lowered(_strip(spec), False, False, spatial)
  oh, ow, is_gaussian = low.program.output_shape(H, W)
  want_ntk = get_c is None or 'ntk' in get_c
  want_cov = get_c is None or 'cov1' in get_c or 'cov2' in get_c
  ctx = _lib.get_context()
  flags = (_lib.FLAG_NO_FUSION if config.disable_fusion else 0) | (_lib.FLAG_PER_LAYER if config.per_layer else 0) | (_lib.FLAG_FULL_SQUARE if config.full_square else 0)
  res = _lib.gram_host(ctx, low.program, x1c, x2c, H, W, C, flags, oh, ow, want_ntk, want_cov)
  m = low.out_meta
  ntk = res['ntk']
  if want_ntk and not m.ntk_tensor:
    ntk = np.zeros((), dt)                                   # requirements.py:807
  cov1, cov2 = res['cov1'], res['cov2']
  shape1 = _out_shape(spec, tuple(x1.shape))
  shape2 = _out_shape(spec, tuple(x1.shape if x2 is None else x2.shape))
  out_spatial = oh > 0
  return Kernel(
      nngp=_to_ref_layout(res['nngp'], m.is_reversed, 2),
      ntk=_to_ref_layout(ntk, m.is_reversed, 2) if ntk is not None and ntk.ndim else ntk,
      cov1=_to_ref_layout(cov1, m.is_reversed, 1),
      cov2=_to_ref_layout(cov2, m.is_reversed, 1),
      x1_is_x2=_x1_is_x2(x1c, x2c), is_gaussian=is_gaussian,
      is_reversed=m.is_reversed if out_spatial else False, is_input=False, diagonal_batch=True,
      diagonal_spatial=False, shape1=tuple(shape1), shape2=tuple(shape2), batch_axis=0,
      channel_axis=len(shape1) - 1, mask1=None, mask2=None)


def _sym_rows(kernel_fn, x, r0, r1, names):
  """Rows [r0, r1) of the symmetric Gram K(x, x), columns [r0, n) only: `{name: [r1 - r0, n - r0]}` with the
  entries (i, j < i) of the leading square unspecified (NTK_FLAG_UPPER_ONLY; the reference computes the full
  square, `_src/batching.py:370`).  `kernel_fn` must come from this module and produce [n1, n2] matrices."""
  spec = kernel_fn._spec
  dt = config.dtype
  xs = np.ascontiguousarray(x[r0:], dt)
  spatial = x.ndim == 4
  H, W = (x.shape[1], x.shape[2]) if spatial else (0, 0)
  low = _lowered(_strip(spec), False, False, spatial)
  oh, ow, _ = low.program.output_shape(H, W)
  if oh > 0:
    raise NotImplementedError('row slabs of a symmetric Gram need [n1, n2] matrix outputs')
  flags = (_lib.FLAG_UPPER_ONLY | (_lib.FLAG_NO_FUSION if config.disable_fusion else 0) |
           (_lib.FLAG_PER_LAYER if config.per_layer else 0))
  res = _lib.gram_host(_lib.get_context(), low.program, xs[:r1 - r0], xs, H, W, x.shape[-1], flags, 0, 0,
                       'ntk' in names, False)
  return {n: res[n] for n in names}


def _gram_on_device(kernel_fn, x1, x2, names):
  """`kernel_fn(x1, x2, names)` with the [n1, n2] results left in HBM: `{name: distributed.DeviceArray}` on this
  thread's GPU (the consumer is `predict`'s device-side Cholesky: no host round trip of the Gram matrices)."""
  from .distributed import DeviceArray
  spec = kernel_fn._spec
  dt = np.dtype(config.dtype)
  _check_inputs(x1, x2)
  spatial = x1.ndim == 4
  H, W = (x1.shape[1], x1.shape[2]) if spatial else (0, 0)
  low = _lowered(_strip(spec), False, False, spatial)
  oh, _, _ = low.program.output_shape(H, W)
  if oh > 0:
    raise NotImplementedError('device-resident Grams are [n1, n2] matrices; this network keeps spatial axes')
  ctx = _lib.get_context()
  n1 = x1.shape[0]
  n2 = n1 if x2 is None else x2.shape[0]
  flags = (_lib.FLAG_NO_FUSION if config.disable_fusion else 0) | (_lib.FLAG_FULL_SQUARE if config.full_square else 0)
  d1 = DeviceArray(ctx, x1.shape, dt)
  d2 = None if x2 is None else DeviceArray(ctx, x2.shape, dt)
  out = {n: DeviceArray(ctx, (n1, n2), dt) for n in ('nngp',) + (('ntk',) if 'ntk' in names else ())}
  try:
    k1 = ctx.h2d(d1.ptr, np.ascontiguousarray(x1, dt))
    k2 = None if x2 is None else ctx.h2d(d2.ptr, np.ascontiguousarray(x2, dt))
    _lib.gram_device(ctx, low.program, dt, d1.ptr, n1, None if d2 is None else d2.ptr, n2, H, W, x1.shape[-1], flags,
                     out['nngp'].ptr, out['ntk'].ptr if 'ntk' in out else None, n2)
    ctx.synchronize()
    del k1, k2
  finally:
    d1.free()
    if d2 is not None:
      d2.free()
  for n in list(out):
    if n not in names:
      out.pop(n).free()
  return out


def _apply_to_kernel(spec, k: Kernel):
  if not k.diagonal_batch:
    raise NotImplementedError('`diagonal_batch=False` kernels are outside the B200 hot path')
  if k.mask1 is not None or k.mask2 is not None:
    raise NotImplementedError('masked kernels are outside the B200 hot path')
  nd = k.nngp.ndim
  if nd == 4 and k.diagonal_spatial:
    raise NotImplementedError('`diagonal_spatial=True` input kernels are outside the B200 hot path')
  if nd not in (2, 6):
    raise NotImplementedError(f'unsupported nngp rank {nd}')
  dt = np.dtype(k.nngp.dtype)
  if dt not in (np.float32, np.float64):
    dt = np.dtype(config.dtype)
  spatial = nd == 6
  rev = bool(k.is_reversed) and spatial
  can = lambda m, b: None if m is None else np.ascontiguousarray(_from_ref_layout(np.asarray(m, dt), rev, b))
  nngp, cov1, cov2 = can(k.nngp, 2), can(k.cov1, 1), can(k.cov2, 1)
  if k.ntk is None:
    ntk, mode = None, _lib.NTK_NONE
  elif np.ndim(k.ntk) == 0:
    ntk, mode = None, _lib.NTK_ZERO
  else:
    ntk, mode = can(k.ntk, 2), _lib.NTK_TENSOR
  H, W = (nngp.shape[2], nngp.shape[4]) if spatial else (0, 0)
  low = _lowered(_strip(spec), rev, mode == _lib.NTK_TENSOR, spatial)
  oh, ow, _ = low.program.output_shape(H, W, k.is_gaussian)
  res = _lib.apply_host(_lib.get_context(), low.program, dt, nngp, ntk, cov1, cov2, H, W, mode,
                        k.is_gaussian, oh, ow)
  m = low.out_meta
  out_spatial = oh > 0
  ntk_o = res['ntk']
  if ntk_o is not None and ntk_o.ndim:
    ntk_o = _to_ref_layout(ntk_o, m.is_reversed, 2)
  shape1 = _out_shape(spec, tuple(k.shape1))
  shape2 = _out_shape(spec, tuple(k.shape2))
  return k.replace(nngp=_to_ref_layout(res['nngp'], m.is_reversed, 2), ntk=ntk_o,
                   cov1=_to_ref_layout(res['cov1'], m.is_reversed, 1),
                   cov2=_to_ref_layout(res['cov2'], m.is_reversed, 1),
                   is_gaussian=res['is_gaussian'],
                   is_reversed=m.is_reversed if out_spatial else False, is_input=False,
                   diagonal_spatial=False, shape1=tuple(shape1), shape2=tuple(shape2),
                   batch_axis=0, channel_axis=len(shape1) - 1)


def _layer(spec, init_fn, apply_fn, req=None):
  """`req`: the static requirements of the layer as the reference's `@requires(...)` states them; `nt.batch` and
  the combinators read them from `kernel_fn.input_req` (`_src/stax/requirements.py:371-386,1053`)."""
  return init_fn, apply_fn, _make_kernel_fn(spec, FrozenDict(req or {}))


def _only_supported(**conds):
  for name, (value, allowed) in conds.items():
    if value not in allowed:
      raise NotImplementedError(f'{name}={value!r} is outside the B200 hot path '
                                f'(supported: {allowed})')


# ----------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------
def Dense(out_dim: int, W_std: float = 1., b_std: Optional[float] = None, batch_axis: int = 0,
          channel_axis: int = -1, parameterization: str = 'ntk', s=(1, 1)):
  """`_src/stax/linear.py:762-926`."""
  _only_supported(parameterization=(parameterization, ('ntk',)), batch_axis=(batch_axis, (0,)),
                  channel_axis=(channel_axis, (-1,)))
  spec = ('dense', float(W_std), None if b_std is None else float(b_std), int(out_dim))

  def init_fn(rng, input_shape):
    g = _rng(rng)
    W = g.standard_normal((input_shape[-1], out_dim))
    b = None if b_std is None else g.standard_normal((out_dim,))
    return tuple(input_shape[:-1]) + (out_dim,), (W, b)

  def apply_fn(params, inputs, **kwargs):
    W, b = params
    out = W_std / math.sqrt(inputs.shape[-1]) * (inputs @ W)
    return out if b is None else out + b_std * b

  return _layer(spec, init_fn, apply_fn, dict(batch_axis=0, channel_axis=-1, diagonal_spatial=Diagonal()))


def Conv(out_chan: int, filter_shape, strides=None, padding: str = 'VALID', W_std: float = 1.,
         b_std: Optional[float] = None, dimension_numbers=None, parameterization: str = 'ntk',
         s=(1, 1)):
  """`_src/stax/linear.py:929-1424` (shared weights, NHWC/HWIO, 2 spatial dims)."""
  _only_supported(parameterization=(parameterization, ('ntk',)))
  if dimension_numbers is not None and tuple(dimension_numbers) != ('NHWC', 'HWIO', 'NHWC'):
    raise NotImplementedError('only NHWC/HWIO/NHWC dimension_numbers are on the B200 hot path')
  if len(filter_shape) != 2:
    raise NotImplementedError('only 2-D convolutions are on the B200 hot path')
  strides = tuple(strides) if strides is not None else (1, 1)
  padding = getattr(padding, 'name', padding).upper()
  if padding not in _lib.PAD:
    raise ValueError(f'unknown padding {padding}')
  k = tuple(int(v) for v in filter_shape)
  spec = ('conv', k, strides, padding, float(W_std), None if b_std is None else float(b_std),
          int(out_chan))

  def init_fn(rng, input_shape):
    g = _rng(rng)
    W = g.standard_normal(k + (input_shape[-1], out_chan))
    b = None if b_std is None else g.standard_normal((1, 1, 1, out_chan))
    return _out_shape(spec, input_shape), (W, b)

  def apply_fn(params, inputs, **kwargs):
    W, b = params
    win = _windows(inputs, k, strides, padding)
    out = np.einsum('nhwabc,abco->nhwo', win, W, optimize=True)
    out = W_std / math.sqrt(inputs.shape[-1] * k[0] * k[1]) * out
    return out if b is None else out + b_std * b

  # shared weights mix positions: the output is never spatially diagonal (`linear.py:1321-1324`)
  return _layer(spec, init_fn, apply_fn, dict(batch_axis=0, channel_axis=3, diagonal_spatial=Diagonal(output=Bool.NO)))


def ABRelu(a: float, b: float, do_stabilize: bool = False):
  """`_src/stax/elementwise.py:403-479`."""
  spec = ('abrelu', float(a), float(b), bool(do_stabilize))
  init_fn = lambda rng, input_shape: (input_shape, ())
  apply_fn = lambda params, inputs, **kw: a * np.minimum(inputs, 0) + b * np.maximum(inputs, 0)
  return _layer(spec, init_fn, apply_fn, dict(diagonal_spatial=Diagonal()))


def Relu(do_stabilize: bool = False):
  """`_src/stax/elementwise.py:482-491`."""
  return ABRelu(0, 1, do_stabilize)


def LeakyRelu(alpha: float, do_stabilize: bool = False):
  """`_src/stax/elementwise.py:494-505`."""
  return ABRelu(alpha, 1, do_stabilize)


def Abs(do_stabilize: bool = False):
  """`_src/stax/elementwise.py:508-517`."""
  return ABRelu(-1, 1, do_stabilize)


def Erf(a: float = 1., b: float = 1., c: float = 0.):
  """`_src/stax/elementwise.py:47-114`."""
  spec = ('erf', float(a), float(b), float(c))
  init_fn = lambda rng, input_shape: (input_shape, ())

  def apply_fn(params, inputs, **kw):
    from scipy.special import erf as _erf
    return a * _erf(b * inputs) + c

  return _layer(spec, init_fn, apply_fn, dict(diagonal_spatial=Diagonal()))


def Sigmoid_like():
  """`_src/stax/elementwise.py:115-126`."""
  return Erf(a=0.5, b=1 / 2.4020563531719796, c=0.5)


def Gelu(approximate: bool = False):
  """`_src/stax/elementwise.py:195-263`; `approximate` only affects the finite-width `apply_fn`."""
  spec = ('gelu',)
  init_fn = lambda rng, input_shape: (input_shape, ())

  def apply_fn(params, inputs, **kw):
    if approximate:
      return 0.5 * inputs * (1 + np.tanh(math.sqrt(2 / math.pi) * (inputs + 0.044715 * inputs**3)))
    from scipy.special import erf as _erf
    return 0.5 * inputs * (1 + _erf(inputs / math.sqrt(2)))

  return _layer(spec, init_fn, apply_fn, dict(diagonal_spatial=Diagonal()))


def Sin(a: float = 1., b: float = 1., c: float = 0.):
  """`_src/stax/elementwise.py:266-320`: `a sin(b x + c)`."""
  spec = ('sin', float(a), float(b), float(c))
  init_fn = lambda rng, input_shape: (input_shape, ())
  apply_fn = lambda params, inputs, **kw: a * np.sin(b * inputs + c)
  return _layer(spec, init_fn, apply_fn, dict(diagonal_spatial=Diagonal()))


def Cos(a: float = 1., b: float = 1., c: float = 0.):
  """`_src/stax/elementwise.py:323-341`: `Sin(a, b, c + pi / 2)`."""
  return Sin(a=a, b=b, c=c + math.pi / 2)


def Rbf(gamma: float = 1.0):
  """`_src/stax/elementwise.py:344-400`: dual activation `sqrt(2) sin(sqrt(2 gamma) x + pi/4)`."""
  spec = ('rbf', float(gamma))
  init_fn = lambda rng, input_shape: (input_shape, ())
  apply_fn = lambda params, inputs, **kw: math.sqrt(2) * np.sin(math.sqrt(2 * gamma) * inputs + math.pi / 4)
  return _layer(spec, init_fn, apply_fn, dict(diagonal_spatial=Diagonal()))


def LayerNorm(axis=-1, eps: float = 1e-12, batch_axis: int = 0, channel_axis: int = -1):
  """`_src/stax/linear.py:2476-2590`, normalisation over the channel axis (the reference's default `axis=-1`;
  normalising over spatial axes as well is outside the B200 hot path)."""
  _only_supported(batch_axis=(batch_axis, (0,)), channel_axis=(channel_axis, (-1,)))
  ax = tuple(axis) if isinstance(axis, (tuple, list)) else (axis,)
  if ax != (-1,):
    raise NotImplementedError('LayerNorm over axes other than the channel axis (-1) is outside the B200 hot path')
  spec = ('layernorm', float(eps))
  init_fn = lambda rng, input_shape: (input_shape, ())

  def apply_fn(params, inputs, **kw):
    mean = inputs.mean(axis=-1, keepdims=True)
    var = inputs.var(axis=-1, keepdims=True)
    return (inputs - mean) / np.sqrt(eps + var)

  return _layer(spec, init_fn, apply_fn, dict(batch_axis=0, channel_axis=-1))


def SumPool(window_shape, strides=None, padding: str = 'VALID', batch_axis: int = 0, channel_axis: int = -1):
  """`_src/stax/linear.py:1503-1547`: AvgPool without the division by the window."""
  _only_supported(batch_axis=(batch_axis, (0,)), channel_axis=(channel_axis, (-1,)))
  if len(window_shape) != 2:
    raise NotImplementedError('only 2-D pooling is on the B200 hot path')
  w = tuple(int(v) for v in window_shape)
  strides = tuple(strides) if strides is not None else (1, 1)
  padding = getattr(padding, 'name', padding).upper()
  if padding not in _lib.PAD:
    raise ValueError(f'unknown padding {padding}')
  spec = ('sumpool', w, strides, padding, False)

  def init_fn(rng, input_shape):
    return _out_shape(spec, input_shape), ()

  def apply_fn(params, inputs, **kwargs):
    return _windows(inputs, w, strides, padding).sum(axis=(3, 4))

  return _layer(spec, init_fn, apply_fn, dict(batch_axis=0, channel_axis=-1, diagonal_spatial=Diagonal(input=Bool.MAYBE)))


def GlobalSumPool(batch_axis: int = 0, channel_axis: int = -1):
  """`_src/stax/linear.py:1674-1720`."""
  _only_supported(batch_axis=(batch_axis, (0,)), channel_axis=(channel_axis, (-1,)))
  spec = ('gsp',)
  init_fn = lambda rng, input_shape: ((input_shape[0], input_shape[-1]), ())
  apply_fn = lambda params, inputs, **kw: inputs.sum(axis=tuple(range(1, inputs.ndim - 1)))
  return _layer(spec, init_fn, apply_fn, dict(batch_axis=0, channel_axis=-1, diagonal_spatial=Diagonal(input=Bool.MAYBE, output=Bool.YES)))


def AvgPool(window_shape, strides=None, padding: str = 'VALID', normalize_edges: bool = False,
            batch_axis: int = 0, channel_axis: int = -1):
  """`_src/stax/linear.py:1459-1501,1565-1671`."""
  _only_supported(batch_axis=(batch_axis, (0,)), channel_axis=(channel_axis, (-1,)))
  if len(window_shape) != 2:
    raise NotImplementedError('only 2-D pooling is on the B200 hot path')
  w = tuple(int(v) for v in window_shape)
  strides = tuple(strides) if strides is not None else (1, 1)
  padding = getattr(padding, 'name', padding).upper()
  if padding not in _lib.PAD:
    raise ValueError(f'unknown padding {padding}')
  spec = ('avgpool', w, strides, padding, bool(normalize_edges))

  def init_fn(rng, input_shape):
    return _out_shape(spec, input_shape), ()

  def apply_fn(params, inputs, **kwargs):
    win = _windows(inputs, w, strides, padding)
    out = win.sum(axis=(3, 4))
    if padding == 'SAME' and normalize_edges:
      cnt = _windows(np.ones_like(inputs), w, strides, padding).sum(axis=(3, 4))
      return out / cnt
    return out / (w[0] * w[1])

  return _layer(spec, init_fn, apply_fn, dict(batch_axis=0, channel_axis=-1, diagonal_spatial=Diagonal(input=Bool.MAYBE)))


def GlobalAvgPool(batch_axis: int = 0, channel_axis: int = -1):
  """`_src/stax/linear.py:1723-1808`."""
  _only_supported(batch_axis=(batch_axis, (0,)), channel_axis=(channel_axis, (-1,)))
  spec = ('gap',)
  init_fn = lambda rng, input_shape: ((input_shape[0], input_shape[-1]), ())
  apply_fn = lambda params, inputs, **kw: inputs.mean(axis=tuple(range(1, inputs.ndim - 1)))
  return _layer(spec, init_fn, apply_fn, dict(batch_axis=0, channel_axis=-1, diagonal_spatial=Diagonal(input=Bool.MAYBE, output=Bool.YES)))


def Flatten(batch_axis: int = 0, batch_axis_out: int = 0):
  """`_src/stax/linear.py:1811-1907`."""
  _only_supported(batch_axis=(batch_axis, (0,)), batch_axis_out=(batch_axis_out, (0,)))
  spec = ('flatten',)
  init_fn = lambda rng, input_shape: ((input_shape[0], int(np.prod(input_shape[1:]))), ())
  apply_fn = lambda params, inputs, **kw: inputs.reshape(inputs.shape[0], -1)
  return _layer(spec, init_fn, apply_fn, dict(batch_axis=0, channel_axis=None, diagonal_spatial=Diagonal(output=Bool.YES)))


def Identity():
  """`_src/stax/linear.py:107-119`."""
  return _layer(('identity',), lambda rng, s: (s, ()), lambda p, x, **kw: x)


def FanOut(num: int):
  """`_src/stax/branching.py:36-52`."""
  return _layer(('fanout', int(num)), lambda rng, s: ([s] * num, ()), lambda p, x, **kw: [x] * num)


def FanInSum():
  """`_src/stax/branching.py:55-117`."""
  warn = ('`FanIn` layers assume independent inputs which is not verified in the code. Please make '
          'sure to have at least one `Dense` / `Conv` layer in each branch.')

  def init_fn(rng, input_shape):
    return input_shape[0], ()

  def apply_fn(params, inputs, **kw):
    return sum(inputs[1:], inputs[0])

  layer = _layer(('faninsum',), init_fn, apply_fn)
  layer[2]._warning = warn
  return layer


def _warn_fan_in(spec):
  """`_src/stax/branching.py:405-410`: every FanIn layer warns that its inputs are assumed independent."""
  import warnings

  def has(sp):
    return sp[0] == 'faninsum' or (sp[0] in ('serial', 'parallel') and any(has(c) for c in sp[1]))
  if has(spec):
    warnings.warn('`FanIn` layers assume independent inputs which is not verified in the code. '
                  'Please make sure to have at least one `Dense` / `Conv` / `GlobalSelfAttention` '
                  'etc. layer in each branch.')


# ----------------------------------------------------------------------------
# combinators
# ----------------------------------------------------------------------------
def serial(*layers):
  """`_src/stax/combinators.py:40-68`."""
  init_fns, apply_fns, kernel_fns = zip(*layers) if layers else ((), (), ())
  spec = ('serial', tuple(k._spec for k in kernel_fns))

  def init_fn(rng, input_shape):
    g = _rng(rng)
    params = []
    for f in init_fns:
      input_shape, p = f(g, input_shape)
      params.append(p)
    return input_shape, params

  def apply_fn(params, inputs, **kwargs):
    for f, p in zip(apply_fns, params):
      inputs = f(p, inputs, **kwargs)
    return inputs

  import operator
  return _layer(spec, init_fn, apply_fn, _fold_input_req(kernel_fns, operator.rshift))


def parallel(*layers):
  """`_src/stax/combinators.py:166-198`."""
  init_fns, apply_fns, kernel_fns = zip(*layers)
  spec = ('parallel', tuple(k._spec for k in kernel_fns))

  def init_fn(rng, input_shape):
    g = _rng(rng)
    res = [f(g, s) for f, s in zip(init_fns, input_shape)]
    return [r[0] for r in res], [r[1] for r in res]

  def apply_fn(params, inputs, **kwargs):
    return [f(p, x, **kwargs) for f, p, x in zip(apply_fns, params, inputs)]

  import operator
  return _layer(spec, init_fn, apply_fn, _fold_input_req(kernel_fns, operator.and_))
