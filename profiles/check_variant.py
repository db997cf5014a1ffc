"""Compares the Myrtle-10 Gram of NTK_B200_PVAR variants of the dominant packed kernel with the default kernel and
with the float64 oracle (fp32 tolerance of north_star: rtol 1e-4).

  python profiles/check_variant.py 10 11 ...      # one JSON line per variant
"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
code = ("import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests', 'golden')!r})\n"
        "import cases, neural_tangents_b200 as nt\n"
        "x1 = np.random.default_rng(1).standard_normal((6, 32, 32, 3)).astype(np.float32)\n"
        "x2 = np.random.default_rng(2).standard_normal((5, 32, 32, 3)).astype(np.float32)\n"
        "_, _, k = cases.build(cases.myrtle(10), nt.stax)\n"
        "a = k(x1, x2, ('nngp', 'ntk')); s = k(x1, None, ('nngp', 'ntk'))\n"
        "np.savez(sys.argv[1], a0=a.nngp, a1=a.ntk, s0=s.nngp, s1=s.ntk)\n")


def run(v):
  path = f'/tmp/pvar_{v}.npz'
  subprocess.run([sys.executable, '-c', code, path], check=True, env=dict(os.environ, NTK_B200_PVAR=str(v)))
  return np.load(path)


def main():
  import cases
  from oracle import ntk_oracle as O
  x1 = np.random.default_rng(1).standard_normal((6, 32, 32, 3)).astype(np.float32)
  x2 = np.random.default_rng(2).standard_normal((5, 32, 32, 3)).astype(np.float32)
  ref = O.kernel_fn(cases.myrtle(10), x1, x2, ('nngp', 'ntk'))
  sref = O.kernel_fn(cases.myrtle(10), x1, None, ('nngp', 'ntk'))
  base = run(0)
  for v in ['0'] + sys.argv[1:]:
    b = run(v)
    rel = lambda p, q: float(np.abs(p / q - 1).max())
    print(json.dumps({
        'variant': v,
        'bit_identical_to_default': bool(np.array_equal(base['a0'], b['a0']) and np.array_equal(base['a1'], b['a1'])),
        'max_rel_vs_default': max(rel(b['a0'], base['a0']), rel(b['a1'], base['a1'])),
        'max_rel_vs_oracle_nngp': rel(b['a0'], ref[0]), 'max_rel_vs_oracle_ntk': rel(b['a1'], ref[1]),
        'duplicate_diag_max_rel_vs_oracle_ntk': rel(np.diag(b['s1']), np.diag(sref[1])),
        'symmetric_offdiag_max_rel_vs_oracle_ntk': rel(b['s1'][~np.eye(6, dtype=bool)], sref[1][~np.eye(6, dtype=bool)])}))


if __name__ == '__main__':
  main()
