import numpy as np, sys
sys.path.insert(0, "tests/golden"); sys.path.insert(0, ".")
import cases, neural_tangents_b200 as nt
from oracle import ntk_oracle as O
for name, spec in (("myrtle10", cases.myrtle(10)), ("myrtle5", cases.myrtle(5))):
  x1 = np.random.default_rng(5).standard_normal((4,32,32,3)).astype(np.float32)
  ref = O.kernel_fn(spec, x1, None, ("nngp","ntk")); _,_,k = cases.build(spec, nt.stax)
  for x64 in (False, True):
    nt.config.update("enable_x64", x64)
    out = k(x1, None, ("nngp","ntk"))
    dup = k(x1, x1.copy(), ("nngp","ntk"))
    d = np.eye(4, dtype=bool)
    print(name, "x64", x64, "diag rel err nngp %.2e ntk %.2e | offdiag nngp %.2e ntk %.2e | x2=x1copy diag ntk %.2e" % (
      np.abs(out.nngp[d]/ref[0][d]-1).max(), np.abs(out.ntk[d]/ref[1][d]-1).max(),
      np.abs(out.nngp[~d]/ref[0][~d]-1).max(), np.abs(out.ntk[~d]/ref[1][~d]-1).max(),
      np.abs(dup.ntk[d]/ref[1][d]-1).max()))
nt.config.update("enable_x64", False)
