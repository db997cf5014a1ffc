"""A/B of the two pipelined input-Gram kernels (cp.async vs TMA + warp specialisation): results and time.

  python profiles/check_gemm_tma.py        # prints one JSON line
"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = ("import sys, time, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests', 'golden')!r})\n"
        "import cases, neural_tangents_b200 as nt\n"
        "from neural_tangents_b200 import _lib, stax\n"
        "_, _, k = cases.build(cases.fcn(3, 2., 0.05), nt.stax)\n"
        "x1 = np.random.default_rng(1).standard_normal((1000, 784)).astype(np.float32)\n"
        "x2 = np.random.default_rng(2).standard_normal((900, 784)).astype(np.float32)\n"
        "a = k(x1, x2, ('nngp', 'ntk')); s = k(x1[:333], None, ('nngp', 'ntk'))\n"
        "ctx = _lib.get_context(); low = stax._lowered(stax._strip(k._spec), False, False, False)\n"
        "n = 8192; x = np.random.default_rng(3).standard_normal((n, 784)).astype(np.float32)\n"
        "d = ctx.malloc(x.nbytes); ctx.h2d(d, x); o1 = ctx.malloc(n * n * 4); o2 = ctx.malloc(n * n * 4)\n"
        "for _ in range(3): _lib.gram_device(ctx, low.program, np.float32, d, n, d + 0, n, 0, 0, 784, 0, o1, o2, n)\n"
        "ctx.synchronize(); e0, e1 = _lib.Event(), _lib.Event(); e0.record(ctx)\n"
        "for _ in range(10): _lib.gram_device(ctx, low.program, np.float32, d, n, d + 0, n, 0, 0, 784, 0, o1, o2, n)\n"
        "e1.record(ctx); ms = e0.elapsed_ms(e1) / 10\n"
        "np.savez(sys.argv[1], a0=a.nngp, a1=a.ntk, s0=s.nngp, s1=s.ntk, ms=ms)\n")
out = {}
for name, env in (('cp_async', {'NTK_B200_GEMM_CPASYNC': '1'}), ('tma', {})):
  path = f'/tmp/gemm_{name}.npz'
  r = subprocess.run([sys.executable, '-c', code, path], env=dict(os.environ, **env), capture_output=True, text=True,
                     timeout=240)
  if r.returncode != 0:
    out[name] = {'error': r.stderr[-800:]}
    continue
  out[name] = np.load(path)
res = {}
for name in out:
  if isinstance(out[name], dict):
    res[name] = out[name]
  else:
    res[name] = {'ms_per_8192x8192_gram_incl_chain': float(out[name]['ms'])}
if all(not isinstance(v, dict) for v in out.values()):
  a, b = out['cp_async'], out['tma']
  res['bit_identical'] = bool(all(np.array_equal(a[k], b[k]) for k in ('a0', 'a1', 's0', 's1')))
  res['max_rel_diff'] = float(max(np.abs(b[k] / a[k] - 1).max() for k in ('a0', 'a1', 's0', 's1')))
print(json.dumps(res))
