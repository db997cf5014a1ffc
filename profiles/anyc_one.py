import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import cases, neural_tangents_b200 as nt
C = int(sys.argv[1])
_, _, k = cases.build(cases.myrtle(10, 'gap'), nt.stax)
x1 = np.random.default_rng(1).standard_normal((96, 32, 32, C)).astype(np.float32)
x2 = np.random.default_rng(2).standard_normal((96, 32, 32, C)).astype(np.float32)
k(x1, x2, ('nngp', 'ntk'))
