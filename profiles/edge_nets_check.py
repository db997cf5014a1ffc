import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import cases, neural_tangents_b200 as nt
from oracle import ntk_oracle as O
V = lambda **kw: cases.conv(pad='VALID', **kw)
nets = {
  'single_stage_c4': (('serial', [cases.conv(W=1.2, b=0.1), cases.RELU, ('gap',), ('dense', 1., 0.1)]), (16, 16, 4)),
  'single_stage_valid': (('serial', [V(W=1.2, b=0.1), cases.RELU, ('gap',)]), (8, 8, 3)),
  'prepass_valid_emb': (('serial', [V(), cases.RELU, V(W=1.1, b=0.2), cases.RELU, ('gap',), ('dense', 1., 0.)]), (12, 12, 2)),
  'prepass_erf_single': (('serial', [cases.conv(), ('erf', 1., 1., 0.), ('gap',)]), (32, 32, 2)),
  'four_layers_c5': (('serial', [cases.conv(), cases.RELU] * 4 + [('gap',)]), (16, 16, 5)),
}
ok = True
for name, (spec, shape) in nets.items():
  _, _, k = cases.build(spec, nt.stax)
  low = nt.stax._lowered(nt.stax._strip(k._spec), False, False, True)
  x1 = np.random.default_rng(5).standard_normal((3,) + shape).astype(np.float32)
  x2 = np.random.default_rng(6).standard_normal((2,) + shape).astype(np.float32)
  ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk')); sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
  for x64, tol in ((False, 1e-4), (True, 1e-10)):
    nt.config.update('enable_x64', x64)
    out = k(x1, x2, ('nngp', 'ntk')); sym = k(x1, None, ('nngp', 'ntk')); only = k(x1, x2, 'nngp')
    e = max(np.abs(out.nngp / ref[0] - 1).max(), np.abs(out.ntk / ref[1] - 1).max(), np.abs(only / ref[0] - 1).max())
    off = ~np.eye(3, dtype=bool)
    es = np.abs(sym.ntk[off] / sref[1][off] - 1).max()
    good = e < tol and es < tol and np.array_equal(sym.ntk, sym.ntk.T)
    ok = ok and good
    print(name, low.program.path(*shape), 'x64' if x64 else 'f32', f'{e:.2e} {es:.2e}', 'OK' if good else 'FAIL')
print('ALL OK' if ok else 'SOME FAILED')
