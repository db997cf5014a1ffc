"""Generates tests/golden/golden.npz by running the REFERENCE's own stax code.

Run in the build container only:  python tests/golden/generate_golden.py

The reference (/root/reference, unmodified, read-only) is executed in float64 on
the NumPy jax shim (oracle/jax_shim, see its README for what is substituted).
For every case in cases.py the reference network is assembled from the spec with
the reference's public layer constructors and `kernel_fn(x1, x2, get)` is
evaluated; outputs are stored under '<case>/<field>'.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import cases  # noqa: E402
import ref_loader  # noqa: E402


build = cases.build


def main():
  warnings.simplefilter('ignore')
  stax = ref_loader.load_reference_stax()
  out = {}
  for name, (spec, _, _, get) in cases.CASES.items():
    x1, x2 = cases.make_inputs(name)
    _, _, kernel_fn = build(spec, stax)
    import jax.numpy as jnp   # the shim: immutable float64 arrays
    x1d = jnp.asarray(x1, np.float64)
    x2d = None if x2 is None else jnp.asarray(x2, np.float64)
    res = kernel_fn(x1d, x2d, get)
    req = getattr(kernel_fn, 'input_req', {})
    d = req.get('diagonal_spatial')
    # the negotiated `diagonal_spatial` requirement of the whole network (requirements.py:425-515): (input, output)
    out[f'{name}/req_diagonal_spatial'] = np.asarray([-1, -1] if d is None or isinstance(d, bool)
                                                      else [int(d.input), int(d.output)])
    out[f'{name}/req_axes'] = np.asarray([-99 if req.get(k) is None else int(req.get(k))
                                          for k in ('batch_axis', 'channel_axis')])
    if get is None:
      for f in ('nngp', 'ntk', 'cov1', 'cov2'):
        v = getattr(res, f)
        if v is not None:
          out[f'{name}/{f}'] = np.asarray(v, np.float64)
      out[f'{name}/is_reversed'] = np.asarray(bool(res.is_reversed))
      out[f'{name}/is_gaussian'] = np.asarray(bool(res.is_gaussian))
      out[f'{name}/shape1'] = np.asarray(res.shape1)
      out[f'{name}/shape2'] = np.asarray(res.shape2)
    else:
      for f in get:
        out[f'{name}/{f}'] = np.asarray(getattr(res, f), np.float64)
    print(name, {k.split('/')[1]: v.shape for k, v in out.items() if k.startswith(name + '/')})
  path = os.path.join(HERE, 'golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
