"""SURVEY §8f row 1 on the device: blocked fp64 Cholesky / cho_solve / K_td alpha on matrices that live in HBM
(`ntk_chol_*`, `ntk_matmul_f64`; neural-tangents_b200/csrc/linalg.cu), against SciPy at 1e-10, and the
`predict.gradient_descent_mse_ensemble(t=None)` means computed without a host round trip of the Gram matrices."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nt():
  import __graft_entry__ as g
  g.build()
  import neural_tangents_b200 as nt
  yield nt
  nt.config.update('enable_x64', False)


def _spd(n, seed, cond=1e3):
  rng = np.random.default_rng(seed)
  q, _ = np.linalg.qr(rng.standard_normal((n, n)))
  lam = np.geomspace(1., 1. / cond, n)
  a = (q * lam) @ q.T
  return (a + a.T) / 2


@pytest.mark.parametrize('n', [1, 5, 63, 64, 65, 200, 777])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_cholesky_and_solve_match_scipy(nt, n, dtype):
  import scipy.linalg
  from neural_tangents_b200 import _lib
  ctx = _lib.get_context()
  a = _spd(n, n).astype(dtype)
  a64 = a.astype(np.float64)
  d = ctx.malloc(a.nbytes)
  ctx.h2d(d, a)
  for diag_reg, absolute in ((0., False), (1e-3, False), (0.25, True)):
    ref = a64 + np.eye(n) * (diag_reg if absolute else diag_reg * np.trace(a64) / n)
    ch = _lib.DeviceCholesky(ctx, dtype, d, n, n, diag_reg, absolute)
    np.testing.assert_allclose(ch.factor(), np.linalg.cholesky(ref), rtol=1e-10, atol=1e-13)
    for nrhs in (1, 10, 130):
      b = np.random.default_rng(nrhs).standard_normal((n, nrhs))
      want = scipy.linalg.cho_solve(scipy.linalg.cho_factor(ref, lower=True), b)
      np.testing.assert_allclose(ch.solve(b), want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
    # K_td alpha on the device: [m, n] (float32 or float64) times [n, nrhs]
    ktd = np.random.default_rng(3).standard_normal((37, n)).astype(dtype)
    dk = ctx.malloc(ktd.nbytes)
    ctx.h2d(dk, ktd)
    x = np.random.default_rng(4).standard_normal((n, 10))
    np.testing.assert_allclose(ch.matmul(dtype, dk, 37, n, x), ktd.astype(np.float64) @ x, rtol=1e-12, atol=1e-12)
    ctx.free(dk)
    ch.close()
  ctx.free(d)


def test_not_positive_definite_is_reported(nt):
  from neural_tangents_b200 import _lib
  ctx = _lib.get_context()
  a = _spd(100, 1)
  a[70, 70] = -1.0
  d = ctx.malloc(a.nbytes)
  ctx.h2d(d, a)
  with pytest.raises(np.linalg.LinAlgError, match='71-th leading minor'):
    _lib.DeviceCholesky(ctx, np.float64, d, 100, 100)
  ctx.free(d)


@pytest.mark.parametrize('x64', [False, True])
def test_ensemble_means_on_device_equal_host_solve(nt, x64):
  """t = None means: Gram -> Cholesky -> K_td alpha in HBM (device_solve) == SciPy on host copies of the Grams."""
  from oracle import ntk_oracle as O
  nt.config.update('enable_x64', x64)
  spec = cases.myrtle(5, 'gap')
  _, _, kernel_fn = cases.build(spec, nt.stax)
  rng = np.random.default_rng(7)
  x_train = rng.standard_normal((70, 32, 32, 3)).astype(np.float32)
  x_test = rng.standard_normal((9, 32, 32, 3)).astype(np.float32)
  y = rng.standard_normal((70, 10))
  dev = nt.predict.gradient_descent_mse_ensemble(kernel_fn, x_train, y, diag_reg=1e-4, device_solve=True)
  host = nt.predict.gradient_descent_mse_ensemble(kernel_fn, x_train, y, diag_reg=1e-4, device_solve=False)
  a = dev(t=None, x_test=x_test, get=('nngp', 'ntk'))
  b = host(t=None, x_test=x_test, get=('nngp', 'ntk'))
  tol = 1e-9 if x64 else 2e-4      # fp32 Grams: the two paths see differently rounded duplicates on the diagonal
  np.testing.assert_allclose(a.nngp, b.nngp, rtol=tol, atol=tol)
  np.testing.assert_allclose(a.ntk, b.ntk, rtol=tol, atol=tol)
  np.testing.assert_allclose(dev(t=None, x_test=x_test, get='ntk'), a.ntk, rtol=0, atol=0)
  np.testing.assert_array_equal(dev(t=None, x_test=None, get='ntk'), y)
  if x64:   # against the float64 oracle on a problem it can hold in host memory (12 x 12 pairs of 32^4 tensors)
    xs, ys = x_train[:12], y[:12]
    small = nt.predict.gradient_descent_mse_ensemble(kernel_fn, xs, ys, diag_reg=1e-4, device_solve=True)
    k_dd = O.kernel_fn(spec, xs, None, ('ntk',))[0]
    k_td = O.kernel_fn(spec, x_test[:3], xs, ('ntk',))[0]
    A = k_dd + 1e-4 * np.trace(k_dd) / 12 * np.eye(12)
    np.testing.assert_allclose(small(t=None, x_test=x_test[:3], get='ntk'), k_td @ np.linalg.solve(A, ys), rtol=1e-7,
                               atol=1e-9)
  nt.config.update('enable_x64', False)


@pytest.mark.parametrize('n', [1, 2, 5, 64, 129, 500])
@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_jacobi_eigh_matches_numpy(nt, n, dtype):
  """`ntk_eigh_*` (parallel cyclic Jacobi in float64 on the device, eigh.cu) against `np.linalg.eigh`: eigenvalues,
  orthogonality, reconstruction and the V^T x / V z / A V z products `predict` uses, incl. odd n (a bye per round)."""
  from neural_tangents_b200 import _lib
  ctx = _lib.get_context()
  a = _spd(n, 100 + n, cond=1e6).astype(dtype)
  a64 = a.astype(np.float64)
  d = ctx.malloc(a.nbytes)
  ctx.h2d(d, a)
  for diag_reg, absolute in ((0., False), (1e-3, False), (0.25, True)):
    ref = a64 + np.eye(n) * (diag_reg if absolute else diag_reg * np.trace(a64) / n)
    w_ref = np.linalg.eigvalsh(ref)
    e = _lib.DeviceEigh(ctx, dtype, d, n, n, diag_reg, absolute)
    scale = np.abs(w_ref).max()
    np.testing.assert_allclose(e.w, w_ref, rtol=0, atol=1e-12 * scale)
    assert np.all(np.diff(e.w) >= 0)
    v = e.vectors()
    np.testing.assert_allclose(v.T @ v, np.eye(n), atol=1e-12)
    np.testing.assert_allclose((v * e.w) @ v.T, ref, atol=1e-11 * scale)
    x = np.random.default_rng(n).standard_normal((n, 7))
    np.testing.assert_allclose(e.project(x), v.T @ x, atol=1e-12 * np.abs(x).max() * max(n, 1) ** 0.5)
    np.testing.assert_allclose(e.expand(x), v @ x, atol=1e-12 * np.abs(x).max() * max(n, 1) ** 0.5)
    ktd = np.random.default_rng(3).standard_normal((11, n)).astype(dtype)
    dk = ctx.malloc(ktd.nbytes)
    ctx.h2d(dk, ktd)
    np.testing.assert_allclose(e.expand_through(dtype, dk, 11, n, x), ktd.astype(np.float64) @ (v @ x), rtol=1e-10,
                               atol=1e-11 * max(n, 1))
    ctx.free(dk)
    assert e.sweeps <= 20 and e.off_over_norm <= 1e-14
    e.close()
  ctx.free(d)


@pytest.mark.parametrize('x64', [False, True])
def test_ensemble_finite_time_means_on_device_equal_host_eigh(nt, x64):
  """Finite-t means: Gram -> Jacobi eigh -> V f(lambda, t) V^T y -> K_td ... with every [n, n] matrix in HBM
  (device_solve) == the host `np.linalg.eigh` path on copies of the same Grams; t -> infinity meets the Cholesky means."""
  nt.config.update('enable_x64', x64)
  spec = cases.myrtle(5, 'gap')
  _, _, kernel_fn = cases.build(spec, nt.stax)
  rng = np.random.default_rng(8)
  x_train = rng.standard_normal((48, 32, 32, 3)).astype(np.float32)
  x_test = rng.standard_normal((7, 32, 32, 3)).astype(np.float32)
  y = rng.standard_normal((48, 3))
  dev = nt.predict.gradient_descent_mse_ensemble(kernel_fn, x_train, y, diag_reg=1e-4, device_solve=True)
  host = nt.predict.gradient_descent_mse_ensemble(kernel_fn, x_train, y, diag_reg=1e-4, device_solve=False)
  ts = np.array([[0., 1.], [10., 1e3]])
  tol = 1e-9 if x64 else 2e-4
  for xt in (x_test, None):
    a = dev(t=ts, x_test=xt, get=('nngp', 'ntk'))
    b = host(t=ts, x_test=xt, get=('nngp', 'ntk'))
    assert a.ntk.shape == b.ntk.shape == (2, 2, 48 if xt is None else 7, 3)
    np.testing.assert_allclose(a.nngp, b.nngp, rtol=tol, atol=tol)
    np.testing.assert_allclose(a.ntk, b.ntk, rtol=tol, atol=tol)
    np.testing.assert_allclose(dev(t=5., x_test=xt, get='ntk'), host(t=5., x_test=xt, get='ntk'), rtol=tol, atol=tol)
  inf = dev(t=None, x_test=x_test, get='ntk')
  np.testing.assert_allclose(dev(t=1e12, x_test=x_test, get='ntk'), inf, rtol=1e-6 if x64 else 2e-3, atol=1e-6 if x64 else 2e-3)
  nt.config.update('enable_x64', False)
