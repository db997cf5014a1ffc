"""Closed-form inference with the NNGP / NTK Gram matrices: the immediate caller of the hot path
(SURVEY §8f row 1).  Mirrors `neural_tangents.predict.gp_inference` (`_src/predict.py:566-750`),
`gradient_descent_mse_ensemble` (`:753-1100`), `gradient_descent_mse` (`:71-279`) and `max_learning_rate`
(`:1103-1150`) for the case the B200 path produces:
`[n1, n2]` kernel matrices (outputs block-diagonal along the logit axis, `trace_axes=(-1,)`).

The Gram matrices come from `kernel_fn` (libntk_b200.so on the GPU).  `gradient_descent_mse_ensemble(t=None)`
means -- the common case, `K_td (K_dd + reg I)^-1 y` -- never leave the GPU: the train-train Gram stays in HBM,
is factorised there in float64 (`ntk_chol_factor`: blocked Cholesky on the fp64 tensor cores), and only the
[n_test, n_out] predictions are copied back (`_DeviceMeans`).  Finite-time means can stay on the device as well
(`device_solve=True`: `ntk_eigh_*`, parallel Jacobi in float64); by default they, and the posterior covariances, use
host float64 linear algebra (Cholesky / one symmetric eigendecomposition), as the reference does with `jax.scipy.linalg`.
"""
import collections
from typing import Callable, Optional

import numpy as np

Gaussian = collections.namedtuple('Gaussian', 'mean covariance')   # `_src/predict.py:552-563`


def _canonicalize_get(get):
  """`utils.canonicalize_get` (`_src/utils/utils.py:139-155`) for the two names used here."""
  if get is None:
    return True, ('nngp', 'ntk')
  if isinstance(get, str):
    get, single = (get,), True
  else:
    get, single = tuple(get), False
  get = tuple(g.lower() for g in get)
  if not get:
    raise ValueError('"get" must be non-empty.')
  if len(set(get)) < len(get):
    raise ValueError('All entries in "get" must be unique. Got {}'.format(get))
  return (False if not single else None), get


def _pack(get_arg, names, values):
  """str -> the value, tuple / None -> a namedtuple over `names` (`utils.get_namedtuple`)."""
  if isinstance(get_arg, str):
    return values[0]
  return collections.namedtuple('Gaussians', names)(*values)


def _attr(k, name):
  """An ndarray stands for whichever kernel is asked; otherwise read the field (`_get_attr`)."""
  if k is None:
    return None
  if isinstance(k, np.ndarray):
    return k
  v = getattr(k, name, None)
  if v is None:
    raise ValueError(f'The kernel `{name}` is required but missing from {type(k).__name__}.')
  return np.asarray(v)


def _as_matrix(k, what):
  k = np.asarray(k, dtype=np.float64)
  if k.ndim != 2:
    raise NotImplementedError(f'{what} must be an [n1, n2] matrix (got shape {k.shape}); kernels with spatial '
                              'or `trace_axes=()` structure are outside the B200 hot path.')
  return k


def _regularize(a: np.ndarray, diag_reg: float, absolute: bool) -> np.ndarray:
  """K + diag_reg * (mean of the diagonal unless absolute) * I   (`_add_diagonal_regularizer`)."""
  n = a.shape[0]
  scale = diag_reg if absolute else diag_reg * np.trace(a) / n
  out = a.copy()
  out[np.diag_indices(n)] += scale
  return out


class _CholSolver:
  """Cached Cholesky factor of a regularised train-train matrix (`_get_cho_solve`)."""

  def __init__(self, k_dd, diag_reg, absolute):
    import scipy.linalg
    self._sl = scipy.linalg
    self.factor = scipy.linalg.cho_factor(_regularize(k_dd, diag_reg, absolute), lower=False)

  def __call__(self, b):
    return self._sl.cho_solve(self.factor, np.asarray(b, dtype=np.float64))


def _check_targets(y_train, trace_axes):
  y = np.asarray(y_train, dtype=np.float64)
  if y.ndim != 2:
    raise NotImplementedError('y_train must be [n_train, n_outputs].')
  ta = tuple(a % y.ndim for a in (trace_axes if isinstance(trace_axes, (tuple, list)) else (trace_axes,)))
  if ta != (1,):
    raise NotImplementedError('only trace_axes=(-1,) (one kernel shared by all outputs) is supported.')
  return y


def gp_inference(k_train_train, y_train, diag_reg: float = 0., diag_reg_absolute_scale: bool = False,
                 trace_axes=(-1,)) -> Callable:
  """Posterior of the NNGP / NTK / NTKGP given train-train kernels (`_src/predict.py:566-750`).

  Returns `predict_fn(get=None, k_test_train=None, k_test_test=None)`; `get` in 'nngp', 'ntk',
  'ntkgp' or a tuple; a mean, or `Gaussian(mean, covariance)` when `k_test_test` is given.
  """
  y = _check_targets(y_train, trace_axes)
  solvers, alphas = {}, {}

  def solver(name):
    if name not in solvers:
      solvers[name] = _CholSolver(_as_matrix(_attr(k_train_train, name), 'k_train_train'), diag_reg,
                                  diag_reg_absolute_scale)
    return solvers[name]

  def k_inv_y(name):
    if name not in alphas:
      alphas[name] = solver(name)(y)
    return alphas[name]

  def predict_fn(get=None, k_test_train=None, k_test_test=None):
    _, names = _canonicalize_get(get)
    values = []
    for g in names:
      if g not in ('nngp', 'ntk', 'ntkgp'):
        raise ValueError(g)
      k = 'ntk' if g == 'ntkgp' else g
      k_dd = _as_matrix(_attr(k_train_train, k), 'k_train_train')
      k_td = None if k_test_train is None else _as_matrix(_attr(k_test_train, k), 'k_test_train')
      mean = y.copy() if k_td is None else k_td @ k_inv_y(k)
      if k_test_test is None:
        values.append(mean)
        continue
      if k_td is None:                      # train set: N(y_train, 0)
        values.append(Gaussian(mean, np.zeros_like(k_dd)))
        continue
      if g == 'ntk' and (isinstance(k_train_train, np.ndarray) or isinstance(k_test_train, np.ndarray)):
        raise ValueError('The NTK posterior covariance on the test set needs both the NTK and the NNGP '
                         'train-train and test-train matrices (namedtuples with `nngp` and `ntk`).')
      init = 'ntk' if g == 'ntkgp' else 'nngp'              # kernel of the wide net at initialisation
      init_td = _as_matrix(_attr(k_test_train, init), 'k_test_train')
      k_tt = _as_matrix(_attr(k_test_test, init), 'k_test_test')
      kinv_init_dt = solver(k)(init_td.T)                      # K_k^-1 K_init(train, test)
      if g in ('nngp', 'ntkgp'):
        cov = k_tt - k_td @ kinv_init_dt
      else:
        # Theta_td Theta^-1 K_dd Theta^-1 Theta_dt - (Theta_td Theta^-1 K_dt + transpose) + K_tt
        w = solver('ntk')(k_td.T)
        nngp_dd = _as_matrix(_attr(k_train_train, 'nngp'), 'k_train_train')
        cross = k_td @ kinv_init_dt
        cov = w.T @ nngp_dd @ w - (cross + cross.T) + k_tt
      values.append(Gaussian(mean, cov))
    return _pack(get, names, values)

  return predict_fn


class _DeviceMeans:
  """Infinite-time means on the device (`_src/predict.py:858-869,936-939` + `gp_inference` mean, `:681-689`):
  K_dd is computed into HBM, (K_dd + reg I) = L L^T and alpha = (K_dd + reg I)^-1 y are cached per kernel name,
  `mean(name, x_test)` = K_td alpha with K_td produced and consumed on the GPU."""

  def __init__(self, kernel_fn, x_train, y, diag_reg, absolute):
    self.kernel_fn, self.x_train, self.y = kernel_fn, x_train, y
    self.diag_reg, self.absolute = diag_reg, absolute
    self.chol, self.alpha, self.eig = {}, {}, {}

  def _factor(self, name):
    from . import _lib, stax
    if name not in self.chol:
      k = stax._gram_on_device(self.kernel_fn, self.x_train, None, (name,))[name]
      try:
        self.chol[name] = _lib.DeviceCholesky(k.ctx, k.dtype, k.ptr, k.shape[0], k.shape[1], self.diag_reg,
                                              self.absolute)
      finally:
        k.free()
      self.alpha[name] = self.chol[name].solve(self.y)
    return self.chol[name]

  def mean(self, name, x_test):
    from . import stax
    ch = self._factor(name)
    k_td = stax._gram_on_device(self.kernel_fn, x_test, self.x_train, (name,))[name]
    try:
      return ch.matmul(k_td.dtype, k_td.ptr, k_td.shape[0], k_td.shape[1], self.alpha[name])
    finally:
      k_td.free()


  def _eigh(self, name):
    """Eigenbasis of the regularised train-train matrix, computed where the Gram kernels left it (`ntk_eigh_*`)."""
    from . import _lib, stax
    if name not in self.eig:
      k = stax._gram_on_device(self.kernel_fn, self.x_train, None, (name,))[name]
      try:
        self.eig[name] = _lib.DeviceEigh(k.ctx, k.dtype, k.ptr, k.shape[0], k.shape[1], self.diag_reg, self.absolute)
      finally:
        k.free()
    return self.eig[name]

  def mean_t(self, name, x_test, ts, norm):
    """Finite-time means (`_src/predict.py:944-1005`): train set V (1 - e^{-lambda t / |y|}) V^T y, test set
    K_td V ((1 - e^{-lambda t / |y|}) / |lambda|) V^T y for every t in `ts` ([T], already times the learning rate).
    The [n, n] matrices (Gram, V) never leave the device; [n, T * out] panels do."""
    from . import stax
    e = self._eigh(name)
    n = e.n
    lam = np.maximum(e.w, 0.)
    coef = -np.expm1(-np.outer(ts, lam) / norm)                    # [T, n]
    if x_test is not None:
      coef = coef / np.abs(e.w)[None, :]
    vty = e.project(self.y)                                         # [n, out]
    z = np.moveaxis(coef[:, :, None] * vty[None], 0, 1).reshape(n, -1)   # [n, T * out]
    if x_test is None:
      out = e.expand(z)
    else:
      k_td = stax._gram_on_device(self.kernel_fn, x_test, self.x_train, (name,))[name]
      try:
        out = e.expand_through(k_td.dtype, k_td.ptr, k_td.shape[0], k_td.shape[1], z)
      finally:
        k_td.free()
    return np.moveaxis(out.reshape(out.shape[0], len(ts), -1), 1, 0)    # [T, m, out]


def gradient_descent_mse_ensemble(kernel_fn, x_train, y_train, learning_rate: float = 1., diag_reg: float = 0.,
                                  diag_reg_absolute_scale: bool = False, trace_axes=(-1,), device_solve=None,
                                  **kernel_fn_train_train_kwargs) -> Callable:
  """Mean and covariance of an infinite ensemble of infinitely wide networks trained on MSE with
  continuous gradient descent for time `t` (`_src/predict.py:753-1100`).

  Returns `predict_fn(t=None, x_test=None, get=None, compute_cov=False)`.  `t=None` is infinite
  time (linear solves); finite `t` (scalar or array) works in the eigenbasis of the regularised
  train-train matrix with the reference's time normalisation `t * learning_rate / y_train.size`.
  """
  y = _check_targets(y_train, trace_axes)
  norm = float(y.size)
  cache, eig, inf = {}, {}, {}
  # `device_solve`: None = infinite-time means on the GPU (Cholesky) whenever `kernel_fn` is a B200 kernel_fn called
  # without extra kwargs, finite times on the host; False = host SciPy everywhere (the round-1 path); True = require
  # the device path, finite-time means included (Jacobi eigh on the device).
  native = hasattr(kernel_fn, '_spec') and not kernel_fn_train_train_kwargs and isinstance(x_train, np.ndarray)
  if device_solve and not native:
    raise ValueError('device_solve=True needs a kernel_fn built by neural_tangents_b200.stax and no extra kwargs')
  dev = _DeviceMeans(kernel_fn, x_train, y, diag_reg, diag_reg_absolute_scale) if (native and device_solve is not False) \
      else None

  def k_train_train(names):
    missing = tuple(n for n in names if n not in cache)
    if missing:
      res = kernel_fn(x_train, None, missing if len(missing) > 1 else missing[0], **kernel_fn_train_train_kwargs)
      if len(missing) == 1:
        cache[missing[0]] = np.asarray(res)
      else:
        for n in missing:
          cache[n] = np.asarray(getattr(res, n))
    K = collections.namedtuple('Kernel', names)
    return K(*(cache[n] for n in names))

  def eigenspace(name):
    if name not in eig:
      k_dd = _regularize(_as_matrix(getattr(k_train_train((name,)), name), 'k_train_train'), diag_reg,
                         diag_reg_absolute_scale)
      eig[name] = np.linalg.eigh(k_dd)
    return eig[name]

  def dependency(names, compute_cov):
    for g in names:
      if g not in ('nngp', 'ntk'):
        raise NotImplementedError('Can only get either "nngp" or "ntk" predictions, got %s.' % g)
    dep = ()
    if 'nngp' in names or ('ntk' in names and compute_cov):
      dep += ('nngp',)
    if 'ntk' in names:
      dep += ('ntk',)
    return dep

  def predict_fn(t=None, x_test=None, get=None, compute_cov: bool = False, **kernel_fn_test_test_kwargs):
    _, names = _canonicalize_get(get)
    dep = dependency(names, compute_cov)
    if dev is not None and t is None and not compute_cov and not kernel_fn_test_test_kwargs:
      # infinite time, means only: Gram -> Cholesky -> K_td alpha, all in HBM
      vals = [y.copy() if x_test is None else dev.mean(g, x_test) for g in names]
      return _pack(get, names, vals)
    if dev is not None and device_solve and t is not None and not compute_cov and not kernel_fn_test_test_kwargs:
      # finite times, means only, on request (device_solve=True): Gram -> Jacobi eigh -> V f(lambda, t) V^T y (-> K_td ...)
      # with every [n, n] matrix in HBM.  Not the default: the HBM-bound Jacobi sweeps (5.3 TB/s effective) take 2x the
      # time of 16-core LAPACK `dsyevd` on a host copy (profiles/eigh_r02.json), and the copy is only n^2 * 4 bytes.
      t_arr = np.asarray(t, dtype=np.float64) * learning_rate
      vals = []
      for g in names:
        m = dev.mean_t(g, x_test, t_arr.reshape(-1), norm)
        vals.append(m.reshape(t_arr.shape + m.shape[1:]))
      return _pack(get, names, vals)
    k_dd = k_train_train(dep)
    kw = dict(kernel_fn_train_train_kwargs)
    kw.update(kernel_fn_test_test_kwargs)
    if x_test is None:
      k_td, nngp_tt = None, (True if compute_cov else None)
    else:
      res = kernel_fn(x_test, x_train, dep if len(dep) > 1 else dep[0], **kw)
      K = collections.namedtuple('Kernel', dep)
      k_td = K(*( [np.asarray(res)] if len(dep) == 1 else [np.asarray(getattr(res, n)) for n in dep]))
      nngp_tt = np.asarray(kernel_fn(x_test, None, 'nngp', **kw)) if compute_cov else None

    if t is None:                                             # infinite time
      if dep not in inf:
        inf[dep] = gp_inference(k_dd, y, diag_reg, diag_reg_absolute_scale, trace_axes)
      # train set with compute_cov: any non-None k_test_test (the posterior there is N(y_train, 0))
      k_tt = np.zeros((1, 1)) if nngp_tt is True else nngp_tt
      return inf[dep](get=get if get is not None else names, k_test_train=k_td, k_test_test=k_tt)

    t_arr = np.asarray(t, dtype=np.float64) * learning_rate
    t_shape = t_arr.shape
    ts = t_arr.reshape(-1)
    values = []
    for g in names:
      evals, evecs = eigenspace(g)
      lam = np.maximum(evals, 0.)
      decay = np.exp(-np.outer(ts, lam) / norm)               # [T, n]  exp(-lambda t / |y|)
      one_minus = -np.expm1(-np.outer(ts, lam) / norm)        # [T, n]  1 - exp(.)
      vty = evecs.T @ y                                       # [n, out]
      if k_td is None:
        mean = np.einsum('ji,ti,ik->tjk', evecs, one_minus, vty)
      else:
        ktd = _as_matrix(getattr(k_td, g), 'k_test_train')
        inv = one_minus / np.abs(evals)[None, :]              # (1 - exp)/|lambda|
        mean = np.einsum('lj,ji,ti,ik->tlk', ktd, evecs, inv, vty)
      mean = mean.reshape(t_shape + mean.shape[1:])
      if nngp_tt is None:
        values.append(mean)
        continue
      nngp_dd = _as_matrix(k_dd.nngp, 'k_train_train')
      if k_td is None:
        if g == 'nngp':
          cov = np.einsum('ji,ti,ki->tjk', evecs, lam[None, :] * decay ** 2, evecs)
        else:
          e = np.einsum('mi,ti,ki->tmk', evecs, decay, evecs)
          cov = np.einsum('tmk,kl,tnl->tmn', e, nngp_dd, e)
      else:
        ktt = _as_matrix(nngp_tt, 'k_test_test')[None]
        if g == 'nngp':
          inv2 = -np.expm1(-2. * np.outer(ts, lam) / norm) / np.abs(evals)[None, :]
          cov = ktt - np.einsum('mj,ji,ti,ki,lk->tml', ktd, evecs, inv2, evecs, ktd)
        else:
          nngp_td = _as_matrix(k_td.nngp, 'k_test_train')
          term1 = np.einsum('mi,ti,ki,lk->tml', evecs, inv, evecs, ktd)          # [T, n_train, n_test]
          term2 = np.einsum('mj,ji,ti,ki,lk->tml', ktd, evecs, inv, evecs, nngp_td)
          term2 = term2 + np.swapaxes(term2, 1, 2)
          cov = np.einsum('tji,jk,tkl->til', term1, nngp_dd, term1) - term2 + ktt
      cov = cov.reshape(t_shape + cov.shape[1:])
      values.append(Gaussian(mean, cov))
    return _pack(get, names, values)

  return predict_fn


def gradient_descent_mse(k_train_train, y_train, learning_rate: float = 1., diag_reg: float = 0.,
                         diag_reg_absolute_scale: bool = False, trace_axes=(-1,)) -> Callable:
  """Function-space gradient descent on MSE for ONE network with kernel `k_train_train`
  (`_src/predict.py:71-279`).

  Returns `predict_fn(t=None, fx_train_0=0., fx_test_0=None, k_test_train=None)` giving the network outputs on
  the train [and test] set at time[s] `t` from their values at t = 0:
    f_train(t) = y + exp(-K t')(f_train(0) - y),   f_test(t) = f_test(0) + K_td K^-1 (I - exp(-K t'))(y - f_train(0)),
  with t' = t * learning_rate / y_train.size and K the regularised train-train matrix.
  """
  y = _check_targets(y_train, trace_axes)
  k_dd = _as_matrix(k_train_train, 'k_train_train')
  norm = float(y.size)
  state = {}

  def solver():
    if 'chol' not in state:
      state['chol'] = _CholSolver(k_dd, diag_reg, diag_reg_absolute_scale)
    return state['chol']

  def eigenspace():
    if 'eig' not in state:
      state['eig'] = np.linalg.eigh(_regularize(k_dd, diag_reg, diag_reg_absolute_scale))
    return state['eig']

  def predict_fn(t=None, fx_train_0=0., fx_test_0=None, k_test_train=None):
    if fx_train_0 is None and fx_test_0 is None:
      raise ValueError('Both `fx_train_0` and `fx_test_0` are `None`, i.e. no predictions will be computed.')
    if fx_test_0 is not None and k_test_train is None:
      raise ValueError('To get predictions on the test set, please provide `k_test_train`.')
    f0 = None if fx_train_0 is None else np.broadcast_to(np.asarray(fx_train_0, dtype=np.float64), y.shape)
    ft0 = None if fx_test_0 is None else np.asarray(fx_test_0, dtype=np.float64)
    k_td = None if k_test_train is None else _as_matrix(k_test_train, 'k_test_train')
    resid = y if f0 is None else y - f0                      # y - f_train(0)

    if t is None:                                             # infinite time
      if ft0 is None:
        return y.copy()
      test = ft0 + k_td @ solver()(resid)
      return test if f0 is None else (y.copy(), test)

    t_arr = np.asarray(t, dtype=np.float64) * learning_rate
    t_shape, ts = t_arr.shape, t_arr.reshape(-1)
    evals, evecs = eigenspace()
    lam = np.maximum(evals, 0.)
    one_minus = -np.expm1(-np.outer(ts, lam) / norm)          # [T, n]
    vtr = evecs.T @ resid                                     # [n, out]
    out = []
    if f0 is not None:
      d_train = np.einsum('ji,ti,ik->tjk', evecs, one_minus, vtr)
      out.append((f0[None] + d_train).reshape(t_shape + y.shape))
    if ft0 is not None:
      d_test = np.einsum('lj,ji,ti,ik->tlk', k_td, evecs, one_minus / np.abs(evals)[None, :], vtr)
      out.append((ft0[None] + d_test).reshape(t_shape + d_test.shape[1:]))
    return out[0] if len(out) == 1 else tuple(out)

  return predict_fn


def max_learning_rate(ntk_train_train, y_train_size: Optional[int] = None, momentum: float = 0.,
                      eps: float = 1e-12) -> float:
  """Largest stable learning rate of (momentum) gradient descent on MSE for an infinitely wide network
  (`_src/predict.py:1103-1150`): 2 (1 + momentum) |y| / (lambda_max(NTK) + eps)."""
  import scipy.linalg
  k = _as_matrix(ntk_train_train, 'ntk_train_train')
  n = k.shape[0]
  factor = n if y_train_size is None else y_train_size
  lam_max = scipy.linalg.eigvalsh(k, subset_by_index=(n - 1, n - 1))[-1]
  return float(2. * (1. + momentum) * factor / (lam_max + eps))
