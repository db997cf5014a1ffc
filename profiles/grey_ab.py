"""Grey 32x32 inputs: pre-pass + packed kernels (default) against the scalar EMB route (NTK_B200_GREY_EMB=1).

  python profiles/grey_ab.py            # one JSON line per route
"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = f"""
import sys, time, json, numpy as np
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests', 'golden')!r})
import cases, neural_tangents_b200 as nt
out = {{}}
for name, spec in (('myrtle10_gap', cases.myrtle(10, 'gap')), ('myrtle5', cases.myrtle(5))):
  _, _, k = cases.build(spec, nt.stax)
  x1 = np.random.default_rng(1).standard_normal((96, 32, 32, 1)).astype(np.float32)
  x2 = np.random.default_rng(2).standard_normal((96, 32, 32, 1)).astype(np.float32)
  k(x1[:8], x2[:8], ('nngp', 'ntk'))
  t0 = time.perf_counter(); r = k(x1, x2, ('nngp', 'ntk')); dt = time.perf_counter() - t0
  out[name] = dict(entries_per_s=96 * 96 / dt, checksum=float(r.ntk.sum()))
print(json.dumps(out))
"""
for env in ({}, {'NTK_B200_GREY_EMB': '1'}):
  r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, **env), capture_output=True, text=True)
  print(json.dumps({'route': 'emb' if env else 'prepass+packed', 'result': json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-500:]}))
