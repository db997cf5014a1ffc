"""`batching.gram_to_disk`: slab-by-slab Gram computation that survives interruption,
exercised on the CPU with a NumPy kernel_fn."""
import collections
import importlib.util
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AK = collections.namedtuple('AnalyticKernel', 'nngp ntk')


def _batching():
  import sys
  sys.path.insert(0, ROOT)
  import neural_tangents_b200.batching as b      # importing the package does not need a GPU
  return b


def make_kernel(log, fail_at=None):
  def kernel_fn(x1, x2=None, get=None):
    x2 = x1 if x2 is None else x2
    log.append(len(x1))
    if fail_at is not None and len(log) == fail_at:
      raise RuntimeError('interrupted')
    nngp = x1 @ x2.T / x1.shape[1]
    ntk = nngp + np.tanh(nngp)
    if get == 'nngp':
      return nngp
    if get == 'ntk':
      return ntk
    return AK(nngp, ntk)
  return kernel_fn


def test_gram_to_disk_resume_and_result(tmp_path):
  b = _batching()
  rng = np.random.default_rng(0)
  x1, x2 = rng.standard_normal((37, 5)), rng.standard_normal((11, 5))
  ref = make_kernel([])(x1, x2, None)
  log = []
  with pytest.raises(RuntimeError):                         # dies while computing the 3rd of 4 slabs
    b.gram_to_disk(make_kernel(log, fail_at=3), x1, x2, ('nngp', 'ntk'), str(tmp_path), block_rows=10)
  assert log == [10, 10, 10]
  assert sorted(f for f in os.listdir(tmp_path) if 'slab' in f) == [
      'nngp.slab000000000.npy', 'nngp.slab000000010.npy', 'ntk.slab000000000.npy', 'ntk.slab000000010.npy']
  log2 = []
  out = b.gram_to_disk(make_kernel(log2), x1, x2, ('nngp', 'ntk'), str(tmp_path), block_rows=10)
  assert log2 == [10, 7]                                    # only the missing slabs are recomputed
  np.testing.assert_array_equal(out.nngp, ref.nngp)
  np.testing.assert_array_equal(out.ntk, ref.ntk)
  assert isinstance(out.ntk, np.memmap) and not out.ntk.flags.writeable
  log3 = []
  again = b.gram_to_disk(make_kernel(log3), x1, x2, ('nngp', 'ntk'), str(tmp_path), block_rows=10)
  assert log3 == [] and np.array_equal(again.nngp, ref.nngp)
  assert json.load(open(tmp_path / 'manifest.json'))['n1'] == 37
  with pytest.raises(ValueError):                           # another computation must not reuse the directory
    b.gram_to_disk(make_kernel([]), x1[:30], x2, ('nngp', 'ntk'), str(tmp_path), block_rows=10)


def test_gram_to_disk_symmetric_single_get(tmp_path):
  b = _batching()
  x = np.random.default_rng(1).standard_normal((16, 4)).astype(np.float32)
  out = b.gram_to_disk(make_kernel([]), x, None, 'ntk', str(tmp_path), block_rows=8)
  ref = make_kernel([])(x, None, 'ntk')
  assert out.dtype == np.float32 and out.shape == (16, 16)
  np.testing.assert_array_equal(out, ref)
  with pytest.raises(ValueError):
    b.gram_to_disk(make_kernel([]), x, None, (), str(tmp_path / 'e'))


def test_gram_to_disk_symmetric_computes_only_the_upper_trapezoids(tmp_path):
  """x2 = None: slab r0 holds columns [r0, n) only (half the work of the reference's full square,
  `_src/batching.py:370`); the assembled matrix is mirrored; ragged last slab."""
  b = _batching()
  x = np.random.default_rng(2).standard_normal((23, 6))
  shapes = []

  def kernel_fn(x1, x2=None, get=None):
    shapes.append((len(x1), len(x2)))
    k = x1 @ x2.T
    # poison what the schedule declares unneeded: the strictly lower part of the leading square
    k[np.tril_indices(len(x1), -1)] = np.nan
    return AK(k, 2 * k)
  out = b.gram_to_disk(kernel_fn, x, None, ('nngp', 'ntk'), str(tmp_path), block_rows=10)
  assert shapes == [(10, 23), (10, 13), (3, 3)]
  np.testing.assert_allclose(out.nngp, x @ x.T, rtol=1e-13)
  np.testing.assert_allclose(out.ntk, 2 * (x @ x.T), rtol=1e-13)
  assert sum(a * c for a, c in shapes) < 0.7 * 23 * 23


def test_gram_to_disk_refuses_other_data_of_the_same_shape(tmp_path):
  """The manifest fingerprints the inputs: re-running in the same directory with different data of identical
  sizes must not silently reuse stale slabs."""
  b = _batching()
  rng = np.random.default_rng(3)
  x1, x2 = rng.standard_normal((8, 3)), rng.standard_normal((4, 3))
  b.gram_to_disk(make_kernel([]), x1, x2, 'nngp', str(tmp_path), block_rows=4)
  m = json.load(open(tmp_path / 'manifest.json'))
  assert len(m['x1_sha1']) == 40 and m['symmetric'] is False
  with pytest.raises(ValueError, match='x1_sha1'):
    b.gram_to_disk(make_kernel([]), x1 + 1.0, x2, 'nngp', str(tmp_path), block_rows=4)
  with pytest.raises(ValueError, match='x2_sha1'):
    b.gram_to_disk(make_kernel([]), x1, x2 * 2.0, 'nngp', str(tmp_path), block_rows=4)
