"""Device-side Cholesky / cho_solve (neural-tangents_b200/csrc/linalg.cu) against SciPy on the host cores.

  python profiles/linalg_bench.py [n ...]      # one JSON line per size
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  import scipy.linalg
  import __graft_entry__ as g
  g.build()
  from neural_tangents_b200 import _lib
  ctx = _lib.get_context()
  for n in [int(a) for a in sys.argv[1:]] or [2048, 4096, 10000]:
    rng = np.random.default_rng(0)
    f = rng.standard_normal((n, 64))
    a = (f @ f.T / 64 + np.eye(n)).astype(np.float32)            # SPD, condition ~ n / 64
    y = rng.standard_normal((n, 10))
    d = ctx.malloc(a.nbytes)
    ctx.h2d(d, a)
    ctx.synchronize()
    _lib.DeviceCholesky(ctx, np.float32, d, min(n, 256), n).close()   # warm-up (module load)
    e0, e1, e2 = _lib.Event(), _lib.Event(), _lib.Event()
    e0.record(ctx)
    ch = _lib.DeviceCholesky.__new__(_lib.DeviceCholesky)
    import ctypes
    ch._lib, ch.ctx, ch.n, ch._h = _lib.load(), ctx, n, ctypes.c_void_p()
    _lib.check(ch._lib.ntk_chol_factor(ctx.handle, 0, ctypes.c_void_p(d), n, n, 1e-6, 0, ctypes.byref(ch._h)))
    e1.record(ctx)
    t0 = time.perf_counter()
    x = ch.solve(y)
    t_solve = time.perf_counter() - t0
    ms_factor = e0.elapsed_ms(e1)
    a64 = a.astype(np.float64)
    a64[np.diag_indices(n)] += 1e-6 * np.trace(a64) / n
    t0 = time.perf_counter()
    cf = scipy.linalg.cho_factor(a64, lower=True)
    t_host_factor = time.perf_counter() - t0
    t0 = time.perf_counter()
    xh = scipy.linalg.cho_solve(cf, y)
    t_host_solve = time.perf_counter() - t0
    print(json.dumps({'n': n, 'device_factor_ms': ms_factor, 'device_factor_tflops_fp64': n**3 / 3 / (ms_factor * 1e-3) / 1e12,
                      'device_solve_10rhs_ms_incl_copies': 1e3 * t_solve, 'host_scipy_factor_ms': 1e3 * t_host_factor,
                      'host_scipy_solve_ms': 1e3 * t_host_solve, 'host_cores': len(os.sched_getaffinity(0)),
                      'max_rel_diff_vs_scipy': float(np.abs(x / xh - 1).max())}))
    ch.close()
    ctx.free(d)


if __name__ == '__main__':
  main()
