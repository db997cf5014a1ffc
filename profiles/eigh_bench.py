"""Device-side symmetric eigendecomposition (neural-tangents_b200/csrc/eigh.cu, parallel cyclic Jacobi) against
`np.linalg.eigh` on the host cores.

  python profiles/eigh_bench.py [n ...]      # one JSON line per size
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  import __graft_entry__ as g
  g.build()
  from neural_tangents_b200 import _lib
  ctx = _lib.get_context()
  for n in [int(a) for a in sys.argv[1:]] or [512, 2048, 4096]:
    rng = np.random.default_rng(0)
    f = rng.standard_normal((n, 64))
    a = (f @ f.T / 64 + np.eye(n)).astype(np.float32)            # SPD, 64 large eigenvalues + a flat tail
    d = ctx.malloc(a.nbytes)
    ctx.h2d(d, a)
    ctx.synchronize()
    _lib.DeviceEigh(ctx, np.float32, d, min(n, 128), n).close()   # warm-up (module load)
    t0 = time.perf_counter()
    e = _lib.DeviceEigh(ctx, np.float32, d, n, n, 1e-6, False)
    t_dev = time.perf_counter() - t0
    a64 = a.astype(np.float64)
    a64[np.diag_indices(n)] += 1e-6 * np.trace(a64) / n
    t0 = time.perf_counter()
    w = np.linalg.eigvalsh(a64) if n > 6000 else np.linalg.eigh(a64)[0]
    t_host = time.perf_counter() - t0
    print(json.dumps({'n': n, 'device_eigh_s': t_dev, 'sweeps': e.sweeps, 'off_over_norm': e.off_over_norm,
                      'rounds_per_s': e.sweeps * (n - 1 + (n & 1)) / t_dev,
                      'effective_GBps': e.sweeps * (n - 1 + (n & 1)) * 32.0 * n * n / t_dev / 1e9,
                      'host_numpy_s': t_host, 'host_fn': 'eigvalsh' if n > 6000 else 'eigh',
                      'host_cores': len(os.sched_getaffinity(0)),
                      'max_abs_eigenvalue_diff_over_scale': float(np.abs(e.w - w).max() / np.abs(w).max())}), flush=True)
    e.close()
    ctx.free(d)


if __name__ == '__main__':
  main()
