"""Host logic of `neural_tangents_b200.predict` (the caller of the Gram path, `_src/predict.py:566-1100`),
checked on the CPU with a synthetic positive-definite `kernel_fn` against direct dense-algebra formulas
(explicit inverses, `scipy.linalg.expm`) and against the identities the reference's own tests use
(`tests/predict_test.py`: t = inf equals t = None, train-set predictions reproduce y_train, t = 0 is the prior)."""
import collections
import importlib.util
import os
import sys

import numpy as np
import pytest
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_predict():
  spec = importlib.util.spec_from_file_location('ntk_b200_predict', os.path.join(ROOT, 'neural-tangents_b200', 'predict.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


predict = _load_predict()
AK = collections.namedtuple('AnalyticKernel', 'nngp ntk')


def kernel_fn(x1, x2=None, get=None):
  """nngp = RBF, ntk = RBF + a linear kernel: both positive definite, ntk >= nngp like a real NTK."""
  x2 = x1 if x2 is None else x2
  d2 = ((x1[:, None, :] - x2[None, :, :]) ** 2).sum(-1)
  nngp = np.exp(-0.5 * d2 / x1.shape[1])
  ntk = nngp + 0.3 * (x1 @ x2.T) / x1.shape[1] + 0.2 * nngp ** 2
  if get == 'nngp':
    return nngp
  if get == 'ntk':
    return ntk
  return AK(nngp, ntk)


@pytest.fixture(scope='module')
def data():
  rng = np.random.default_rng(0)
  x_train, x_test = rng.standard_normal((24, 6)), rng.standard_normal((7, 6))
  y_train = rng.standard_normal((24, 3))
  return x_train, y_train, x_test


def test_gp_inference_matches_direct_formulas(data):
  x, y, xt = data
  reg = 1e-3
  k_dd, k_td, k_tt = kernel_fn(x), kernel_fn(xt, x), kernel_fn(xt)
  fn = predict.gp_inference(k_dd, y, diag_reg=reg)
  out = fn(get=('nngp', 'ntk', 'ntkgp'), k_test_train=k_td, k_test_test=k_tt)
  n = len(x)
  K = k_dd.nngp + reg * np.trace(k_dd.nngp) / n * np.eye(n)
  T = k_dd.ntk + reg * np.trace(k_dd.ntk) / n * np.eye(n)
  Ki, Ti = np.linalg.inv(K), np.linalg.inv(T)
  np.testing.assert_allclose(out.nngp.mean, k_td.nngp @ Ki @ y, rtol=1e-9, atol=1e-12)
  np.testing.assert_allclose(out.ntk.mean, k_td.ntk @ Ti @ y, rtol=1e-9, atol=1e-12)
  np.testing.assert_allclose(out.nngp.covariance, k_tt.nngp - k_td.nngp @ Ki @ k_td.nngp.T, rtol=1e-8, atol=1e-11)
  np.testing.assert_allclose(out.ntkgp.covariance, k_tt.ntk - k_td.ntk @ Ti @ k_td.ntk.T, rtol=1e-8, atol=1e-11)
  a = k_td.ntk @ Ti
  ntk_cov = a @ k_dd.nngp @ a.T - (a @ k_td.nngp.T + k_td.nngp @ a.T) + k_tt.nngp
  np.testing.assert_allclose(out.ntk.covariance, ntk_cov, rtol=1e-8, atol=1e-11)
  for g in out:
    assert np.linalg.eigvalsh((g.covariance + g.covariance.T) / 2).min() > -1e-9
  # train set: N(y_train, 0); a single str `get` returns the bare value; absolute regularisation
  tr = fn(get='ntk', k_test_train=None, k_test_test=True)
  np.testing.assert_array_equal(tr.mean, y)
  assert not tr.covariance.any()
  m = predict.gp_inference(k_dd, y, diag_reg=0.1, diag_reg_absolute_scale=True)('nngp', k_td)
  np.testing.assert_allclose(m, k_td.nngp @ np.linalg.inv(k_dd.nngp + 0.1 * np.eye(n)) @ y, rtol=1e-9)
  with pytest.raises(ValueError):                      # NTK covariance needs nngp and ntk
    predict.gp_inference(k_dd.ntk, y)(get='ntk', k_test_train=k_td.ntk, k_test_test=k_tt.nngp)
  with pytest.raises(NotImplementedError):
    predict.gp_inference(k_dd, y, trace_axes=())


def test_ensemble_finite_time_against_matrix_exponential(data):
  x, y, xt = data
  fn = predict.gradient_descent_mse_ensemble(kernel_fn, x, y, learning_rate=0.7, diag_reg=1e-4)
  ts = np.array([[0.0, 3.0], [50.0, 4000.0]])
  out = fn(t=ts, x_test=xt, get=('nngp', 'ntk'), compute_cov=True)
  k_dd, k_td, k_tt = kernel_fn(x), kernel_fn(xt, x), kernel_fn(xt)
  n = len(x)
  for g in ('nngp', 'ntk'):
    A = getattr(k_dd, g) + 1e-4 * np.trace(getattr(k_dd, g)) / n * np.eye(n)
    Ai = np.linalg.inv(A)
    res = getattr(out, g)
    assert res.mean.shape == (2, 2, 7, 3) and res.covariance.shape == (2, 2, 7, 7)
    for idx in np.ndindex(ts.shape):
      tau = ts[idx] * 0.7 / y.size
      E = scipy.linalg.expm(-A * tau)
      np.testing.assert_allclose(res.mean[idx], getattr(k_td, g) @ Ai @ (np.eye(n) - E) @ y, rtol=1e-7, atol=1e-10)
      if g == 'nngp':
        cov = k_tt.nngp - k_td.nngp @ Ai @ (np.eye(n) - scipy.linalg.expm(-2 * A * tau)) @ k_td.nngp.T
      else:
        B = Ai @ (np.eye(n) - E)
        cov = k_td.ntk @ B @ k_dd.nngp @ B.T @ k_td.ntk.T - (k_td.ntk @ B @ k_td.nngp.T + k_td.nngp @ B.T @ k_td.ntk.T) + k_tt.nngp
      np.testing.assert_allclose(res.covariance[idx], cov, rtol=1e-6, atol=1e-9)
  # t = 0: the prior (zero mean, NNGP covariance); t -> inf equals t = None
  np.testing.assert_allclose(out.ntk.mean[0, 0], 0., atol=1e-12)
  np.testing.assert_allclose(out.ntk.covariance[0, 0], k_tt.nngp, rtol=1e-10)
  inf = fn(t=None, x_test=xt, get=('nngp', 'ntk'), compute_cov=True)
  far = fn(t=1e12, x_test=xt, get=('nngp', 'ntk'), compute_cov=True)
  for g in ('nngp', 'ntk'):
    np.testing.assert_allclose(getattr(far, g).mean, getattr(inf, g).mean, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(getattr(far, g).covariance, getattr(inf, g).covariance, rtol=1e-5, atol=1e-8)


def test_ensemble_train_set_and_get_conventions(data):
  x, y, xt = data
  calls = []

  def counting(x1, x2=None, get=None):
    calls.append((len(x1), None if x2 is None else len(x2), get))
    return kernel_fn(x1, x2, get)

  fn = predict.gradient_descent_mse_ensemble(counting, x, y)
  # no regularisation: the train set is interpolated at t = None, and approached at large finite t
  np.testing.assert_array_equal(fn(t=None, x_test=None, get='ntk'), y)
  np.testing.assert_allclose(fn(t=1e9, x_test=None, get='ntk'), y, rtol=1e-4, atol=1e-5)
  tr = fn(t=5.0, x_test=None, get='nngp', compute_cov=True)
  assert tr.mean.shape == y.shape and tr.covariance.shape == (len(x), len(x))
  m = fn(x_test=xt, get='ntk')                           # mean only: no test-test kernel is computed
  assert m.shape == (7, 3)
  assert not any(c[1] is None and c[0] == 7 for c in calls)
  n_train_train = sum(1 for c in calls if c[0] == len(x) and c[1] is None)
  fn(x_test=xt, get=('ntk',))
  assert sum(1 for c in calls if c[0] == len(x) and c[1] is None) == n_train_train   # train-train kernels are cached
  both = fn(x_test=xt)                                   # get=None -> namedtuple(nngp, ntk)
  assert both._fields == ('nngp', 'ntk')
  with pytest.raises(NotImplementedError):
    fn(x_test=xt, get='ntkgp')
  with pytest.raises(ValueError):
    fn(x_test=xt, get=('ntk', 'ntk'))


def test_gradient_descent_mse_and_max_learning_rate(data):
  x, y, xt = data
  rng = np.random.default_rng(3)
  k_dd, k_td = kernel_fn(x, None, 'ntk'), kernel_fn(xt, x, 'ntk')
  f0, ft0 = 0.1 * rng.standard_normal(y.shape), 0.1 * rng.standard_normal((len(xt), y.shape[1]))
  fn = predict.gradient_descent_mse(k_dd, y, learning_rate=0.5, diag_reg=1e-5)
  n = len(x)
  A = k_dd + 1e-5 * np.trace(k_dd) / n * np.eye(n)
  ts = np.array([0., 2., 300.])
  tr, te = fn(ts, f0, ft0, k_td)
  assert tr.shape == (3,) + y.shape and te.shape == (3,) + ft0.shape
  for i, t in enumerate(ts):
    E = scipy.linalg.expm(-A * t * 0.5 / y.size)
    np.testing.assert_allclose(tr[i], y + E @ (f0 - y), rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(te[i], ft0 + k_td @ np.linalg.solve(A, (np.eye(n) - E) @ (y - f0)), rtol=1e-7, atol=1e-10)
  np.testing.assert_allclose(tr[0], f0, atol=1e-12)                        # t = 0: nothing has moved
  tr_inf, te_inf = fn(None, f0, ft0, k_td)
  np.testing.assert_array_equal(tr_inf, y)
  np.testing.assert_allclose(fn(1e13, f0, ft0, k_td)[1], te_inf, rtol=1e-6, atol=1e-9)
  assert fn(5., None, ft0, k_td).shape == ft0.shape                        # test only, default f_train(0) = 0 path
  assert fn(5.).shape == y.shape
  with pytest.raises(ValueError):
    fn(1., 0., ft0, None)
  with pytest.raises(ValueError):
    fn(1., None, None)
  lam = np.linalg.eigvalsh(k_dd)[-1]
  assert predict.max_learning_rate(k_dd, y.size, momentum=0.9) == pytest.approx(2 * 1.9 * y.size / (lam + 1e-12))
  assert predict.max_learning_rate(k_dd) == pytest.approx(2 * n / (lam + 1e-12))
