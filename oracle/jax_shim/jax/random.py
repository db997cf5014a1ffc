"""Only what import-time / apply_fn code touches; kernel_fn never draws."""
import numpy as _np


def PRNGKey(seed):
  return _np.array([0, seed], dtype=_np.uint32)


def split(key, num=2):
  return [_np.array([int(key[0]) + i + 1, int(key[1])], dtype=_np.uint32) for i in range(num)]


def fold_in(key, data):
  return _np.array([int(key[0]) ^ (data + 0x9E37), int(key[1])], dtype=_np.uint32)


def _rng(key):
  return _np.random.default_rng([int(key[0]), int(key[1])])


def normal(key, shape=(), dtype=_np.float64):
  return _rng(key).standard_normal(shape).astype(dtype)


def bernoulli(key, p=0.5, shape=()):
  return _rng(key).random(shape) < p


def randint(key, shape, minval, maxval):
  return _rng(key).integers(minval, maxval, shape)
