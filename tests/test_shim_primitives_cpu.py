"""Pins the L0 primitives of `oracle/jax_shim` against an independent implementation (PyTorch on the CPU).

The golden vectors are produced by running the reference's own covariance-space formulas on top of these
NumPy stand-ins for `lax.conv_general_dilated`, `lax.reduce_window`, `lax.dot_general` and `jnp.tensordot`
(tests/golden/generate_golden.py); the formulas are the reference's, the primitives are ours.  These tests close
that last link on the exact call shapes the reference issues:
  * `_src/stax/linear.py:3167-3172`  depthwise NCHW/OIHW conv, rhs = diag(1/k) broadcast over the channels,
                                     equal strides on both axes, 'SAME' / 'VALID' padding names;
  * `_src/stax/linear.py:3554`       6-D additive reduce_window with unit windows on the batch axes;
  * `_src/stax/requirements.py:534`  dot_general contracting the channel axis with a batch axis;
  * `_src/stax/requirements.py:548`  tensordot over the channel axis.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, 'oracle', 'jax_shim')


@pytest.fixture(scope='module')
def shim():
  if SHIM not in sys.path:
    sys.path.insert(0, SHIM)
  import jax.lax as lax
  import jax.numpy as jnp
  return lax, jnp


def _same_pads(n, k, s):
  out = -(-n // s)
  tot = max((out - 1) * s + k - n, 0)
  return tot // 2, tot - tot // 2


def _torch_conv(lhs, rhs, strides, padding, groups):
  import torch
  import torch.nn.functional as F
  x = torch.from_numpy(np.ascontiguousarray(lhs))
  w = torch.from_numpy(np.ascontiguousarray(rhs))
  if padding == 'SAME':
    (t, b), (l, r) = (_same_pads(x.shape[2], w.shape[2], strides[0]), _same_pads(x.shape[3], w.shape[3], strides[1]))
    x = F.pad(x, (l, r, t, b))
  return F.conv2d(x, w, stride=strides, groups=groups).numpy()


@pytest.mark.parametrize('padding', ['SAME', 'VALID'])
@pytest.mark.parametrize('k,s,hw', [(3, 1, (8, 8)), (3, 2, (8, 8)), (3, 2, (7, 7)), (2, 1, (6, 6)), (4, 3, (11, 11))])
def test_depthwise_diag_conv_as_issued_by_the_conv_kernel_rule(shim, padding, k, s, hw):
  """linear.py:3341-3378: lhs [prod(pre), C, n, n'] with rhs [C, 1, k, k] = diag(1/k), strides (s, s)."""
  lax, jnp = shim
  rng = np.random.default_rng(0)
  C = 5
  lhs = rng.standard_normal((3, C, hw[0], hw[1]))
  rhs = np.broadcast_to(np.diag(np.full((k,), 1. / k)), (C, 1, k, k)).copy()
  got = np.asarray(lax.conv_general_dilated(lhs, rhs, (s, s), padding, dimension_numbers=('NCHW', 'OIHW', 'NCHW'),
                                            feature_group_count=C))
  want = _torch_conv(lhs, rhs, (s, s), padding, C)
  assert got.shape == want.shape
  np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize('groups', [1, 2, 6])
def test_grouped_conv_random_weights_and_unequal_strides(shim, groups):
  lax, jnp = shim
  rng = np.random.default_rng(1)
  lhs = rng.standard_normal((2, 6, 9, 10))
  rhs = rng.standard_normal((12, 6 // groups, 3, 2))
  for padding in ('SAME', 'VALID'):
    got = np.asarray(lax.conv_general_dilated(lhs, rhs, (2, 1), padding, feature_group_count=groups))
    want = _torch_conv(lhs, rhs, (2, 1), padding, groups)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)


def test_conv_dimension_numbers_nhwc_hwio(shim):
  """The finite-width `stax.Conv` of the shim's example_libraries uses NHWC / HWIO."""
  lax, jnp = shim
  rng = np.random.default_rng(2)
  x = rng.standard_normal((2, 7, 6, 3))
  w = rng.standard_normal((3, 3, 3, 4))
  got = np.asarray(lax.conv_general_dilated(x, w, (1, 1), 'SAME', dimension_numbers=('NHWC', 'HWIO', 'NHWC')))
  want = _torch_conv(x.transpose(0, 3, 1, 2), w.transpose(3, 2, 0, 1), (1, 1), 'SAME', 1).transpose(0, 2, 3, 1)
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)


def _box_sum_axis(x, axis, k, s, padding):
  """Independent 1-D additive window along `axis` through torch.conv1d with a ones kernel."""
  import torch
  import torch.nn.functional as F
  xm = np.moveaxis(x, axis, -1)
  shp = xm.shape
  t = torch.from_numpy(np.ascontiguousarray(xm)).reshape(-1, 1, shp[-1])
  if padding == 'SAME':
    lo, hi = _same_pads(shp[-1], k, s)
    t = F.pad(t, (lo, hi))
  out = F.conv1d(t, torch.ones(1, 1, k, dtype=t.dtype), stride=s).numpy()
  return np.moveaxis(out.reshape(shp[:-1] + (out.shape[-1],)), -1, axis)


@pytest.mark.parametrize('padding', ['VALID', 'SAME'])
@pytest.mark.parametrize('win,stride', [((2, 2), (2, 2)), ((2, 2), (1, 1)), ((3, 2), (2, 1))])
def test_reduce_window_6d_as_issued_by_the_pool_rule(shim, padding, win, stride):
  """linear.py:3499-3559: window (1, 1, wh, wh, ww, ww), strides (1, 1, sh, sh, sw, sw) on [n1, n2, H, H, W, W]."""
  lax, jnp = shim
  x = np.random.default_rng(3).standard_normal((2, 3, 6, 6, 5, 5))
  w6 = (1, 1, win[0], win[0], win[1], win[1])
  s6 = (1, 1, stride[0], stride[0], stride[1], stride[1])
  got = np.asarray(lax.reduce_window(x, 0., lax.add, w6, s6, padding))
  want = x
  for axis in range(2, 6):
    want = _box_sum_axis(want, axis, w6[axis], s6[axis], padding)
  assert got.shape == want.shape
  np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
  if padding == 'VALID' and win == stride == (2, 2):          # non-overlapping: plain reshape-sum
    r = x[:, :, :6, :6, :4, :4].reshape(2, 3, 3, 2, 3, 2, 2, 2, 2, 2).sum(axis=(3, 5, 7, 9))
    np.testing.assert_allclose(got, r, rtol=1e-12, atol=1e-13)


def test_dot_general_and_tensordot_input_covariances(shim):
  """requirements.py:529-553: cov of x with itself per sample (batch axis) and the x1 x x2 cross covariance."""
  import torch
  lax, jnp = shim
  rng = np.random.default_rng(4)
  x1 = rng.standard_normal((3, 4, 5, 2))
  x2 = rng.standard_normal((2, 4, 5, 2))
  got = np.asarray(lax.dot_general(x1, x1, (((3,), (3,)), ((0,), (0,)))))
  want = torch.einsum('nhwc,nijc->nhwij', torch.from_numpy(x1), torch.from_numpy(x1)).numpy()
  np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-14)
  got = np.asarray(jnp.tensordot(x1, x2, (3, 3)))
  want = torch.tensordot(torch.from_numpy(x1), torch.from_numpy(x2), dims=([3], [3])).numpy()
  np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-14)


def test_padtype_to_pads_matches_lax_rule(shim):
  """SAME: out = ceil(n / s), total = max((out - 1) s + k - n, 0), low = total // 2 (the extra pad goes high)."""
  lax, jnp = shim
  assert lax.padtype_to_pads((32, 7), (3, 3), (1, 2), 'SAME') == [(1, 1), (1, 1)]
  assert lax.padtype_to_pads((32, 8), (3, 3), (2, 2), 'SAME') == [(0, 1), (0, 1)]
  assert lax.padtype_to_pads((6,), (2,), (1,), 'SAME') == [(0, 1)]
  assert lax.padtype_to_pads((6, 6), (3, 3), (1, 1), 'VALID') == [(0, 0), (0, 0)]
