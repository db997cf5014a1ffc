"""Bit-compares the Myrtle-10 Gram of a NTK_B200_PVAR variant of the dominant packed kernel with the default."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = ("import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests', 'golden')!r})\n"
        "import cases, neural_tangents_b200 as nt\n"
        "x1 = np.random.default_rng(1).standard_normal((6, 32, 32, 3)).astype(np.float32)\n"
        "x2 = np.random.default_rng(2).standard_normal((5, 32, 32, 3)).astype(np.float32)\n"
        "_, _, k = cases.build(cases.myrtle(10), nt.stax)\n"
        "a = k(x1, x2, ('nngp', 'ntk')); np.savez(sys.argv[1], a0=a.nngp, a1=a.ntk)\n")
res = {}
for v in ('0', sys.argv[1]):
  path = f'/tmp/pvar_{v}.npz'
  subprocess.run([sys.executable, '-c', code, path], check=True, env=dict(os.environ, NTK_B200_PVAR=v))
  res[v] = np.load(path)
a, b = res['0'], res[sys.argv[1]]
print('variant', sys.argv[1], 'bit-identical:', bool(np.array_equal(a['a0'], b['a0']) and np.array_equal(a['a1'], b['a1'])),
      'max rel diff', float(max(np.abs(b['a0'] / a['a0'] - 1).max(), np.abs(b['a1'] / a['a1'] - 1).max())))
