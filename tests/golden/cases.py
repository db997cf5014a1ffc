"""Golden-vector case list shared by the generator (reference side) and the tests.

Each case is plain data: a network *spec* (see oracle/ntk_oracle.py docstring),
input shapes and RNG seeds.  Inputs are regenerated from the seeds with
`numpy.random.default_rng(seed).standard_normal(shape)`, so the
committed fixture only needs the reference outputs.
"""
import numpy as np

SQ2 = 2 ** 0.5
RELU = ('abrelu', 0., 1., False)


def conv(k=(3, 3), s=(1, 1), pad='SAME', W=SQ2, b=0.):
  return ('conv', tuple(k), tuple(s), pad, W, b)


def pool(w=(2, 2), s=(2, 2), pad='VALID', ne=False):
  return ('avgpool', tuple(w), tuple(s), pad, ne)


def myrtle(depth, tail='flatten'):
  f = {5: [2, 1, 1], 7: [2, 2, 2], 10: [3, 3, 3]}[depth]
  layers = [conv(), RELU] * f[0] + [pool()] + [conv(), RELU] * f[1] + [pool()] + [conv(), RELU] * f[2]
  layers += ([pool()] * 3 + [('flatten',)]) if tail == 'flatten' else [('gap',)]
  layers += [('dense', SQ2, 0.)]
  return ('serial', layers)


def fcn(depth=3, W=2., b=0.05):
  layers = []
  for _ in range(depth):
    layers += [('dense', W, b), RELU]
  return ('serial', layers + [('dense', W, b)])


def wrn_block(stride, channel_mismatch, act=RELU):
  """README.md:192-222 WideResnetBlock: pre-activation residual block."""
  main = ('serial', [act, conv(s=(stride, stride), W=1., b=0.1), act, conv(W=1., b=0.1)])
  shortcut = ('identity',) if not channel_mismatch else conv(s=(stride, stride), W=1., b=0.1)
  return ('serial', [('fanout', 2), ('parallel', [main, shortcut]), ('faninsum',)])


def wrn(act=RELU):
  return ('serial', [conv(W=1., b=0.1),
                     wrn_block(1, True, act), wrn_block(1, False, act),
                     wrn_block(2, True, act), wrn_block(1, False, act),
                     ('gap',), ('dense', 1., 0.1)])


# name -> (spec, x1_shape, x2_shape_or_None, get)
CASES = {
    'fcn3': (fcn(3), (6, 784), (5, 784), ('nngp', 'ntk')),
    'fcn3_sym': (fcn(3), (7, 784), None, ('nngp', 'ntk')),
    'fcn1_nngp': (fcn(1), (4, 32), (3, 32), ('nngp',)),
    'myrtle5': (myrtle(5), (2, 32, 32, 3), (2, 32, 32, 3), ('nngp', 'ntk')),
    'myrtle5_gap': (myrtle(5, 'gap'), (2, 32, 32, 3), (1, 32, 32, 3), ('nngp', 'ntk')),
    'myrtle7': (myrtle(7), (2, 32, 32, 3), (1, 32, 32, 3), ('nngp', 'ntk')),
    'myrtle10': (myrtle(10), (2, 32, 32, 3), (2, 32, 32, 3), ('nngp', 'ntk')),
    'myrtle10_sym': (myrtle(10), (2, 32, 32, 3), None, ('nngp', 'ntk')),
    'myrtle10_16px': (('serial', [conv(), RELU] * 3 + [pool()] + [conv(), RELU] * 3 + [pool()] +
                       [conv(), RELU] * 3 + [('gap',), ('dense', SQ2, 0.)]),
                      (3, 16, 16, 3), (4, 16, 16, 3), ('nngp', 'ntk')),
    'conv_pool_stride': (('serial', [conv(W=1.3, b=0.2), RELU, pool(), conv(s=(2, 2), W=1.1, b=0.1),
                                     RELU, ('gap',), ('dense', 1.2, 0.3)]),
                         (3, 8, 8, 3), (4, 8, 8, 3), ('nngp', 'ntk')),
    'conv_flatten_bias': (('serial', [conv(W=1.5, b=0.3), RELU, conv(W=1.2, b=0.1), RELU,
                                      ('flatten',), ('dense', 1., 0.5)]),
                          (3, 6, 5, 2), (2, 6, 5, 2), ('nngp', 'ntk')),
    'erf_valid': (('serial', [conv(W=1.2, b=0.1), ('erf', 1., 1., 0.),
                              conv(k=(2, 3), pad='VALID', W=1.1, b=0.2), ('erf', 0.8, 1.3, 0.2),
                              ('flatten',), ('dense', 1., 0.1)]),
                  (3, 6, 5, 2), (4, 6, 5, 2), ('nngp', 'ntk')),
    'erf_pool_same': (('serial', [conv(W=1.2, b=0.1), ('erf', 1., 1., 0.),
                                  pool((2, 2), (1, 1), 'SAME', False), conv(W=1., b=0.),
                                  ('erf', 1., 0.7, 0.), ('gap',), ('dense', 1., 0.)]),
                      (2, 6, 6, 3), (3, 6, 6, 3), ('nngp', 'ntk')),
    'circular': (('serial', [conv(pad='CIRCULAR', W=1.3, b=0.1), RELU,
                             pool((2, 2), (2, 2), 'CIRCULAR', False),
                             conv(k=(3, 2), pad='CIRCULAR'), RELU, ('gap',), ('dense', 1., 0.)]),
                 (2, 8, 6, 3), (3, 8, 6, 3), ('nngp', 'ntk')),
    'pool_norm_edges': (('serial', [conv(), RELU, pool((3, 3), (2, 2), 'SAME', True), conv(), RELU,
                                    ('gap',), ('dense', 1., 0.)]),
                        (2, 7, 7, 2), (2, 7, 7, 2), ('nngp', 'ntk')),
    'leaky_abs': (('serial', [conv(W=1.1, b=0.2), ('abrelu', 0.2, 1., False), conv(),
                              ('abrelu', -1., 1., False), ('gap',), ('dense', 1., 0.1)]),
                  (3, 5, 5, 2), (2, 5, 5, 2), ('nngp', 'ntk')),
    'relu_stabilize': (('serial', [conv(W=1.1, b=0.2), ('abrelu', 0., 1., True), conv(),
                                   ('abrelu', 0., 1., True), ('gap',), ('dense', 1., 0.1)]),
                       (3, 5, 5, 2), (2, 5, 5, 2), ('nngp', 'ntk')),
    'wrn_relu': (wrn(RELU), (3, 8, 8, 3), (2, 8, 8, 3), ('nngp', 'ntk')),
    'wrn_erf': (wrn(('erf', 1., 1., 0.)), (2, 8, 8, 3), (3, 8, 8, 3), ('nngp', 'ntk')),
    'wrn_relu_sym': (wrn(RELU), (3, 8, 8, 3), None, ('nngp', 'ntk')),
    # Full `Kernel` outputs (get=None) with spatial axes: layout / is_reversed contract (F4).
    'kernel_conv_relu': (('serial', [conv(W=1.2, b=0.1), RELU]), (2, 4, 3, 2), (3, 4, 3, 2), None),
    'kernel_conv2_pool': (('serial', [conv(W=1.2, b=0.1), RELU, conv(k=(3, 2)), RELU, pool((2, 1), (2, 1))]),
                          (2, 4, 4, 2), None, None),
    # ---- round 2 (append-only: the input seeds derive from the case order) ----
    # MNIST geometry: 28x28x1 through the Myrtle-5 body (28 -> 14 -> 7) with a GlobalAvgPool tail
    'myrtle5_gap_mnist': (myrtle(5, 'gap'), (2, 28, 28, 1), (3, 28, 28, 1), ('nngp', 'ntk')),
    'myrtle5_gap_mnist_sym': (myrtle(5, 'gap'), (3, 28, 28, 1), None, ('nngp', 'ntk')),
    # grey 32x32, non-square RGB, odd size with a VALID pool (15 -> 7)
    'myrtle5_grey32': (myrtle(5), (2, 32, 32, 1), (2, 32, 32, 1), ('nngp', 'ntk')),
    'conv_pool_20x12': (('serial', [conv(W=1.3, b=0.1), RELU, conv(), RELU, pool(), conv(W=1.1, b=0.2), RELU,
                                    ('gap',), ('dense', 1.2, 0.1)]),
                        (2, 20, 12, 3), (3, 20, 12, 3), ('nngp', 'ntk')),
    'conv_pool_15_odd': (('serial', [conv(W=1.3, b=0.1), RELU, pool(), conv(), ('abrelu', 0.1, 1., False),
                                     ('gap',), ('dense', 1., 0.)]),
                         (2, 15, 15, 1), (2, 15, 15, 1), ('nngp', 'ntk')),
    # SumPool / GlobalSumPool (linear.py:1503, 1674)
    'sumpool_gsp': (('serial', [conv(W=1.2, b=0.1), RELU, ('sumpool', (2, 2), (2, 2), 'VALID'), conv(), RELU,
                                ('gsp',), ('dense', 1., 0.1)]),
                    (2, 8, 8, 3), (3, 8, 8, 3), ('nngp', 'ntk')),
    'sumpool_same_stride1': (('serial', [conv(W=1.2, b=0.1), ('erf', 1., 1., 0.), ('sumpool', (3, 2), (1, 1), 'SAME'),
                                         ('gap',), ('dense', 1., 0.)]),
                             (2, 5, 6, 2), (2, 5, 6, 2), ('nngp', 'ntk')),
    # closed-form activations of elementwise.py:195-400
    'gelu_conv': (('serial', [conv(W=1.2, b=0.1), ('gelu',), conv(W=1.1, b=0.), ('gelu',), ('gap',),
                              ('dense', 1., 0.1)]),
                  (3, 6, 6, 2), (2, 6, 6, 2), ('nngp', 'ntk')),
    'gelu_fcn_sym': (('serial', [('dense', 1.5, 0.1), ('gelu',), ('dense', 1.2, 0.05), ('gelu',), ('dense', 1., 0.)]),
                     (5, 16), None, ('nngp', 'ntk')),
    'sin_cos_conv': (('serial', [conv(W=1.1, b=0.2), ('sin', 1.2, 0.7, 0.3), conv(W=1., b=0.1), ('cos', 0.9, 1.1, 0.2),
                                 ('flatten',), ('dense', 1., 0.1)]),
                     (2, 5, 4, 3), (3, 5, 4, 3), ('nngp', 'ntk')),
    'rbf_fcn': (('serial', [('dense', 1., 0.), ('rbf', 0.7), ('dense', 1.3, 0.1), ('rbf', 1.5), ('dense', 1., 0.)]),
                (4, 10), (3, 10), ('nngp', 'ntk')),
    'rbf_conv_pool': (('serial', [conv(W=1., b=0.1), ('rbf', 0.5), pool(), conv(), RELU, ('gap',), ('dense', 1., 0.)]),
                      (2, 6, 6, 2), (2, 6, 6, 2), ('nngp', 'ntk')),
    # LayerNorm over the channel axis (linear.py:2476)
    'layernorm_conv': (('serial', [conv(W=1.3, b=0.2), ('layernorm', 1e-12), RELU, conv(W=1., b=0.1), ('layernorm', 0.1),
                                   ('gelu',), ('gap',), ('dense', 1., 0.1)]),
                       (2, 5, 5, 2), (3, 5, 5, 2), ('nngp', 'ntk')),
    'layernorm_fcn_sym': (('serial', [('dense', 1.5, 0.3), ('layernorm', 1e-6), RELU, ('dense', 1., 0.)]),
                          (4, 12), None, ('nngp', 'ntk')),
    # 3x3 / 1 / VALID convs (linear.py:3341-3378 without padding): Flatten tail (diagonal-column kernels), and pools + a
    # GlobalAvgPool tail (stage kernels: 16 -> 12 -pool-> 6 -> 4, the box origin is even at the pool)
    'valid_flatten': (('serial', [conv(pad='VALID', W=1.3, b=0.1), RELU, conv(pad='VALID'), ('abrelu', 0.1, 1., False),
                                  conv(pad='VALID', W=1.1, b=0.05), RELU, ('flatten',), ('dense', 1., 0.1)]),
                      (3, 12, 12, 3), (2, 12, 12, 3), ('nngp', 'ntk')),
    'valid_pool_gap': (('serial', [conv(pad='VALID', W=1.3, b=0.1), RELU, conv(pad='VALID'), RELU, pool(),
                                   conv(pad='VALID', W=1.1, b=0.2), ('abrelu', 0.2, 1., False), ('gap',),
                                   ('dense', 1.2, 0.1)]),
                       (2, 16, 16, 3), (3, 16, 16, 3), ('nngp', 'ntk')),
    'valid_pool_gap_mnist_sym': (('serial', [conv(pad='VALID'), RELU, conv(pad='VALID'), RELU, pool(),
                                             conv(pad='VALID'), RELU, conv(pad='VALID'), RELU, ('gap',), ('dense', SQ2, 0.)]),
                                 (3, 28, 28, 1), None, ('nngp', 'ntk')),
    # 3x3 / 1 / CIRCULAR convs ending in Flatten (diagonal-column kernels, taps mod S), Relu and Erf
    'circular_flatten': (('serial', [conv(pad='CIRCULAR', W=1.3, b=0.1), RELU, conv(pad='CIRCULAR'), ('erf', 1., 0.9, 0.1),
                                     conv(pad='CIRCULAR', W=1.1, b=0.05), RELU, ('flatten',), ('dense', 1., 0.1)]),
                         (3, 9, 9, 2), (2, 9, 9, 2), ('nngp', 'ntk')),
}


def build(spec, stax):
  """Assembles `spec` with the layer constructors of `stax` (the reference's module in the
  generator, `neural_tangents_b200.stax` in the parity tests)."""
  kind = spec[0]
  if kind == 'serial':
    return stax.serial(*[build(s, stax) for s in spec[1]])
  if kind == 'parallel':
    return stax.parallel(*[build(s, stax) for s in spec[1]])
  if kind == 'fanout':
    return stax.FanOut(spec[1])
  if kind == 'faninsum':
    return stax.FanInSum()
  if kind == 'identity':
    return stax.Identity()
  if kind == 'dense':
    return stax.Dense(1, W_std=spec[1], b_std=spec[2])
  if kind == 'conv':
    return stax.Conv(1, spec[1], strides=spec[2], padding=spec[3], W_std=spec[4], b_std=spec[5])
  if kind == 'abrelu':
    return stax.ABRelu(spec[1], spec[2], do_stabilize=spec[3])
  if kind == 'erf':
    return stax.Erf(spec[1], spec[2], spec[3])
  if kind == 'avgpool':
    return stax.AvgPool(spec[1], strides=spec[2], padding=spec[3], normalize_edges=spec[4])
  if kind == 'gap':
    return stax.GlobalAvgPool()
  if kind == 'sumpool':
    return stax.SumPool(spec[1], strides=spec[2], padding=spec[3])
  if kind == 'gsp':
    return stax.GlobalSumPool()
  if kind == 'gelu':
    return stax.Gelu()
  if kind == 'sin':
    return stax.Sin(spec[1], spec[2], spec[3])
  if kind == 'cos':
    return stax.Cos(spec[1], spec[2], spec[3])
  if kind == 'rbf':
    return stax.Rbf(spec[1])
  if kind == 'layernorm':
    return stax.LayerNorm(eps=spec[1])
  if kind == 'flatten':
    return stax.Flatten()
  raise ValueError(kind)


def make_inputs(name):
  _, s1, s2, _ = CASES[name]
  idx = list(CASES).index(name)  # seeds are derived from the (append-only) case order
  x1 = np.random.default_rng(1000 + 2 * idx).standard_normal(s1).astype(np.float32)
  x2 = None if s2 is None else np.random.default_rng(1001 + 2 * idx).standard_normal(s2).astype(np.float32)
  return x1, x2
