"""NumPy finite-width `stax` layers (init_fun, apply_fun), following the public
`jax.example_libraries.stax` API semantics.  On the golden-generation path they
are used only for output-shape inference (`requirements.py:833-879`), which the
reference runs through `eval_shape`.
"""
import functools
import operator as op

import numpy as _np

from .. import lax, random


def Dense(out_dim, W_init=None, b_init=None):
  def init_fun(rng, input_shape):
    output_shape = tuple(input_shape[:-1]) + (out_dim,)
    k1, k2 = random.split(rng)
    W = W_init(k1, (input_shape[-1], out_dim))
    b = b_init(k2, (out_dim,))
    return output_shape, (W, b)

  def apply_fun(params, inputs, **kwargs):
    W, b = params
    return inputs @ W + b
  return init_fun, apply_fun


def _conv_out_shape(input_shape, kernel_shape, strides, padding, dimension_numbers):
  lhs_spec, rhs_spec, out_spec = dimension_numbers
  sp = [c for c in rhs_spec if c not in 'OI']
  in_sp = [input_shape[lhs_spec.index(c)] for c in sp]
  k_sp = [kernel_shape[rhs_spec.index(c)] for c in sp]
  pads = lax.padtype_to_pads(in_sp, k_sp, strides, padding)
  out_sp = [(n + lo + hi - k) // s + 1 for n, k, s, (lo, hi) in zip(in_sp, k_sp, strides, pads)]
  vals = {'N': input_shape[lhs_spec.index('N')], 'C': kernel_shape[rhs_spec.index('O')]}
  vals.update(dict(zip(sp, out_sp)))
  return tuple(vals[c] for c in out_spec)


def GeneralConv(dimension_numbers, out_chan, filter_shape, strides=None,
                padding='VALID', W_init=None, b_init=None):
  lhs_spec, rhs_spec, out_spec = dimension_numbers
  one = (1,) * len(filter_shape)
  strides = strides or one

  def init_fun(rng, input_shape):
    it = iter(filter_shape)
    kernel_shape = [out_chan if c == 'O' else
                    input_shape[lhs_spec.index('C')] if c == 'I' else next(it)
                    for c in rhs_spec]
    output_shape = _conv_out_shape(input_shape, kernel_shape, strides, padding,
                                   dimension_numbers)
    bias_shape = [out_chan if c == 'C' else 1 for c in out_spec]
    k1, k2 = random.split(rng)
    return output_shape, (W_init(k1, kernel_shape), b_init(k2, bias_shape))

  def apply_fun(params, inputs, **kwargs):
    W, b = params
    return lax.conv_general_dilated(inputs, W, strides, padding,
                                    dimension_numbers=dimension_numbers) + b
  return init_fun, apply_fun


def GeneralConvTranspose(*a, **k):
  def _ni(*a, **k):
    raise NotImplementedError('conv transpose is outside the hot path')
  return _ni, _ni


def _pooling_layer(reducer, init_val, rescaler=None):
  def PoolingLayer(window_shape, strides=None, padding='VALID', spec=None):
    strides_ = strides or (1,) * len(window_shape)
    rescale = rescaler(window_shape, strides_, padding) if rescaler else None
    if spec is None:
      non_spatial = (0, len(window_shape) + 1)
    else:
      non_spatial = (spec.index('N'), spec.index('C'))
    dims, strd = tuple(window_shape), tuple(strides_)
    for i in sorted(non_spatial):
      dims = dims[:i] + (1,) + dims[i:]
      strd = strd[:i] + (1,) + strd[i:]

    def init_fun(rng, input_shape):
      pads = lax.padtype_to_pads(input_shape, dims, strd, padding)
      out_shape = tuple((n + lo + hi - k) // s + 1
                        for n, k, s, (lo, hi) in zip(input_shape, dims, strd, pads))
      return out_shape, ()

    def apply_fun(params, inputs, **kwargs):
      out = lax.reduce_window(inputs, init_val, reducer, dims, strd, padding)
      return rescale(out, inputs, spec) if rescale else out
    return init_fun, apply_fun
  return PoolingLayer


SumPool = _pooling_layer(lax.add, 0.)


def _normalize_by_window_size(dims, strides, padding):
  def rescale(outputs, inputs, spec):
    if not spec:
      non_spatial = (0, inputs.ndim - 1)
    else:
      non_spatial = (spec.index('N'), spec.index('C'))
    spatial_shape = tuple(inputs.shape[i] for i in range(inputs.ndim) if i not in non_spatial)
    one = _np.ones(spatial_shape, dtype=inputs.dtype)
    window_sizes = lax.reduce_window(one, 0., lax.add, dims, strides, padding)
    for i in sorted(non_spatial):
      window_sizes = _np.expand_dims(window_sizes, i)
    return outputs / window_sizes
  return rescale


AvgPool = _pooling_layer(lax.add, 0., _normalize_by_window_size)


def _flatten():
  def init_fun(rng, input_shape):
    return (input_shape[0], functools.reduce(op.mul, input_shape[1:], 1)), ()

  def apply_fun(params, inputs, **kwargs):
    return _np.reshape(inputs, (inputs.shape[0], -1))
  return init_fun, apply_fun


Flatten = _flatten()


def _identity():
  return (lambda rng, input_shape: (input_shape, ())), (lambda params, inputs, **kw: inputs)


Identity = _identity()


def FanOut(num):
  return (lambda rng, input_shape: ([input_shape] * num, ())), \
         (lambda params, inputs, **kw: [inputs] * num)


def _fan_in_sum():
  return (lambda rng, input_shape: (input_shape[0], ())), \
         (lambda params, inputs, **kw: sum(inputs))


FanInSum = _fan_in_sum()


def FanInConcat(axis=-1):
  def _ni(*a, **k):
    raise NotImplementedError
  return _ni, _ni


def Dropout(rate, mode='train'):
  def _ni(*a, **k):
    raise NotImplementedError
  return _ni, _ni


def softmax(*a, **k):
  raise NotImplementedError


def serial(*layers):
  n = len(layers)
  init_funs, apply_funs = zip(*layers) if layers else ((), ())

  def init_fun(rng, input_shape):
    params = []
    for f in init_funs:
      rng, layer_rng = random.split(rng)
      input_shape, p = f(layer_rng, input_shape)
      params.append(p)
    return input_shape, params

  def apply_fun(params, inputs, **kwargs):
    rng = kwargs.pop('rng', None)
    rngs = random.split(rng, n) if rng is not None else (None,) * n
    for f, p, r in zip(apply_funs, params, rngs):
      inputs = f(p, inputs, rng=r, **kwargs)
    return inputs
  return init_fun, apply_fun


def parallel(*layers):
  n = len(layers)
  init_funs, apply_funs = zip(*layers)

  def init_fun(rng, input_shape):
    rngs = random.split(rng, n)
    res = [f(r, s) for f, r, s in zip(init_funs, rngs, input_shape)]
    return [r[0] for r in res], [r[1] for r in res]

  def apply_fun(params, inputs, **kwargs):
    rng = kwargs.pop('rng', None)
    rngs = random.split(rng, n) if rng is not None else (None,) * n
    return [f(p, x, rng=r, **kwargs) for f, p, x, r in zip(apply_funs, params, inputs, rngs)]
  return init_fun, apply_fun
