"""GPU parity: the CUDA path (through the Python front end and the C-ABI) against
(1) the golden vectors generated from the reference's own code, and (2) the oracle on
other seeded inputs.  Tolerances are north_star's: rtol 1e-4 (FP32 path vs the float64
result) and rtol 1e-10 (FP64 path).  Exact-duplicate pairs (the diagonal of a symmetric
Gram) get the documented looser bound (SURVEY Appendix A caveat)."""
import numpy as np
import pytest

import cases
from cases import pool

pytestmark = pytest.mark.gpu

RTOL = {False: 1e-4, True: 1e-10}
# kappa-dot on exact duplicates carries ~sqrt(eps)/pi per Relu layer
RTOL_DUP = {False: 2e-3, True: 2e-7}


@pytest.fixture(scope='module')
def nt():
  import __graft_entry__ as g
  g.build()
  import neural_tangents_b200 as nt
  yield nt
  nt.config.update('enable_x64', False)
  nt.config.update('disable_fusion', False)


def _compare(v, ref, x64, symmetric):
  assert v.shape == ref.shape
  if symmetric and v.ndim == 2:
    off = ~np.eye(v.shape[0], dtype=bool)
    np.testing.assert_allclose(v[off], ref[off], rtol=RTOL[x64], atol=0)
    np.testing.assert_allclose(np.diag(v), np.diag(ref), rtol=RTOL_DUP[x64], atol=0)
  else:
    np.testing.assert_allclose(v, ref, rtol=RTOL[x64], atol=RTOL[x64] * 1e-3)


@pytest.mark.parametrize('fusion', [True, False])
@pytest.mark.parametrize('x64', [False, True])
@pytest.mark.parametrize('name', [n for n, c in cases.CASES.items() if c[3] is not None])
def test_golden_matrix_outputs(nt, golden, name, x64, fusion):
  spec, _, _, get = cases.CASES[name]
  x1, x2 = cases.make_inputs(name)
  nt.config.update('enable_x64', x64)
  nt.config.update('disable_fusion', not fusion)
  _, _, kernel_fn = cases.build(spec, nt.stax)
  try:
    out = kernel_fn(x1, x2, get)
  finally:
    nt.config.update('disable_fusion', False)   # later tests (any -k selection, any order) start on the fused paths
  assert out._fields == tuple(get)
  for f in get:
    v = getattr(out, f)
    assert v.dtype == (np.float64 if x64 else np.float32)
    _compare(v, golden[f'{name}/{f}'], x64, x2 is None)


@pytest.mark.parametrize('x64', [False, True])
@pytest.mark.parametrize('name', [n for n, c in cases.CASES.items() if c[3] is None])
def test_golden_full_kernel_layout(nt, golden, name, x64):
  """get=None returns a `Kernel` in the reference's zipped / is_reversed layout (F4)."""
  spec, _, _, _ = cases.CASES[name]
  x1, x2 = cases.make_inputs(name)
  nt.config.update('enable_x64', x64)
  _, _, kernel_fn = cases.build(spec, nt.stax)
  k = kernel_fn(x1, x2)
  assert isinstance(k, nt.Kernel)
  assert k.is_reversed == bool(golden[f'{name}/is_reversed'])
  assert k.is_gaussian == bool(golden[f'{name}/is_gaussian'])
  assert tuple(k.shape1) == tuple(golden[f'{name}/shape1'])
  assert tuple(k.shape2) == tuple(golden[f'{name}/shape2'])
  for f in ('nngp', 'ntk', 'cov1', 'cov2'):
    key = f'{name}/{f}'
    v = getattr(k, f)
    if key not in golden.files:
      assert v is None
    else:
      # x2=None kernels contain exact-duplicate entries (n1 == n2, h == h', w == w')
      rt = RTOL_DUP[x64] if x2 is None else RTOL[x64]
      np.testing.assert_allclose(v, golden[key], rtol=rt, atol=rt * 1e-3)


@pytest.mark.parametrize('x64', [False, True])
def test_composition_identity(nt, x64):
  """tests/stax/stax_test.py:636-687: kernel_fn(kernel_fn(x)) == composed network."""
  stax = nt.stax
  nt.config.update('enable_x64', x64)
  rng = np.random.default_rng(5)
  x1 = rng.standard_normal((2, 4, 5, 2)).astype(np.float32)
  x2 = rng.standard_normal((3, 4, 5, 2)).astype(np.float32)
  a = stax.serial(stax.Conv(1, (3, 3), padding='SAME', W_std=1.2, b_std=0.1), stax.Relu())
  b = stax.serial(stax.Conv(1, (2, 2), W_std=1.0, b_std=0.3), stax.Relu(), stax.GlobalAvgPool(),
                  stax.Dense(1, 1.1, 0.2))
  ab = stax.serial(a, b)
  k_mid = a[2](x1, x2)
  assert k_mid.is_reversed and k_mid.nngp.shape == (2, 3, 5, 5, 4, 4)
  two = b[2](k_mid)
  one = ab[2](x1, x2)
  tol = 1e-12 if x64 else 1e-5
  np.testing.assert_allclose(two.nngp, one.nngp, rtol=tol)
  np.testing.assert_allclose(two.ntk, one.ntk, rtol=tol)
  np.testing.assert_allclose(two.cov1, one.cov1, rtol=tol)
  assert two.shape1 == one.shape1 == (2, 1)
  # get on a Kernel input
  np.testing.assert_allclose(b[2](k_mid, get='ntk'), one.ntk, rtol=tol)


def test_oracle_parity_other_seeds(nt):
  """CUDA vs oracle on fresh seeds and ragged block shapes (n1 != n2, odd counts)."""
  from oracle import ntk_oracle as O
  spec = cases.myrtle(7, 'gap')
  x1 = np.random.default_rng(11).standard_normal((3, 32, 32, 3)).astype(np.float32)
  x2 = np.random.default_rng(12).standard_normal((5, 32, 32, 3)).astype(np.float32)
  ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
  _, _, kernel_fn = cases.build(spec, nt.stax)
  for x64 in (False, True):
    nt.config.update('enable_x64', x64)
    out = kernel_fn(x1, x2, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
    np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
    only = kernel_fn(x1, x2, 'nngp')                       # nngp-only call skips the ntk work
    np.testing.assert_allclose(only, ref[0], rtol=RTOL[x64])


def test_symmetry_and_x2_none_equivalence(nt):
  """kernel_fn(x, None) == kernel_fn(x, x) and is symmetric (size-independent property)."""
  nt.config.update('enable_x64', False)
  _, _, kernel_fn = cases.build(cases.myrtle(5), nt.stax)
  x = np.random.default_rng(21).standard_normal((6, 32, 32, 3)).astype(np.float32)
  a = kernel_fn(x, None, ('nngp', 'ntk'))
  b = kernel_fn(x, x.copy(), ('nngp', 'ntk'))
  np.testing.assert_allclose(a.nngp, b.nngp, rtol=1e-6)
  np.testing.assert_allclose(a.ntk, b.ntk, rtol=1e-6)
  np.testing.assert_allclose(a.nngp, a.nngp.T, rtol=1e-5)
  np.testing.assert_allclose(a.ntk, a.ntk.T, rtol=1e-5)


def test_batch_equals_unbatched(nt):
  """tests/batching_test.py:160-284: batched == unbatched, incl. divisibility errors."""
  nt.config.update('enable_x64', False)
  _, _, kernel_fn = cases.build(cases.CASES['conv_pool_stride'][0], nt.stax)
  x1 = np.random.default_rng(31).standard_normal((8, 8, 8, 3)).astype(np.float32)
  x2 = np.random.default_rng(32).standard_normal((4, 8, 8, 3)).astype(np.float32)
  full = kernel_fn(x1, x2, ('nngp', 'ntk'))
  for bs in (2, 4):
    b = nt.batch(kernel_fn, batch_size=bs, device_count=0)
    out = b(x1, x2, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, full.nngp, rtol=1e-6)
    np.testing.assert_allclose(out.ntk, full.ntk, rtol=1e-6)
  with pytest.raises(ValueError, match='must divide batch size'):
    nt.batch(kernel_fn, batch_size=3, device_count=0)(x1, x2, 'nngp')   # gcd 4 -> 3; 8 % 3 != 0
  # Kernel outputs through the Python block loop (cov1/cov2 stitched, cov2 None for x2=None)
  _, _, kf_sp = cases.build(('serial', [cases.conv(), cases.RELU]), nt.stax)
  k_full = kf_sp(x1, None)
  k_b = nt.batch(kf_sp, batch_size=4, device_count=0)(x1, None)
  assert k_b.cov2 is None and k_b.shape1 == k_full.shape1
  np.testing.assert_allclose(k_b.nngp, k_full.nngp, rtol=1e-6)
  np.testing.assert_allclose(k_b.cov1, k_full.cov1, rtol=1e-6)


def test_internal_tiling_small_workspace(nt):
  """The executor tiles the pair grid to fit its workspace; results do not depend on it."""
  from neural_tangents_b200 import _lib, stax
  nt.config.update('enable_x64', False)
  spec = cases.CASES['conv_pool_stride'][0]
  _, _, kernel_fn = cases.build(spec, nt.stax)
  x1 = np.random.default_rng(41).standard_normal((7, 8, 8, 3)).astype(np.float32)
  x2 = np.random.default_rng(42).standard_normal((5, 8, 8, 3)).astype(np.float32)
  full = kernel_fn(x1, x2, ('nngp', 'ntk'))
  low = stax._lowered(stax._strip(kernel_fn._spec), False, False, True)
  small = _lib.Context(0, 2 << 20)   # 2 MiB: forces several tiles on the per-layer path
  res = _lib.gram_host(small, low.program, x1, x2, 8, 8, 3, _lib.FLAG_NO_FUSION, 0, 0, True, False)
  np.testing.assert_allclose(res['nngp'], full.nngp, rtol=1e-6)
  np.testing.assert_allclose(res['ntk'], full.ntk, rtol=1e-6)
  small.close()


def test_reference_errors_on_gpu(nt):
  stax = nt.stax
  x = np.random.default_rng(0).standard_normal((2, 4, 4, 2)).astype(np.float32)
  with pytest.raises(ValueError, match='must be Gaussian'):
    stax.serial(stax.Relu())[2](x, None, 'nngp')
  bad = stax.serial(stax.Conv(1, (3, 3), padding='SAME'), stax.FanOut(2),
                    stax.parallel(stax.Relu(), stax.Identity()), stax.FanInSum())
  with pytest.raises(NotImplementedError, match='FanInSum'):      # branching.py:77-85
    bad[2](x, None, 'nngp')
  mism = stax.serial(stax.Conv(1, (3, 3), padding='SAME'), stax.FanOut(2),
                     stax.parallel(stax.Conv(1, (3, 3), padding='VALID'), stax.Identity()),
                     stax.FanInSum())
  with pytest.raises(ValueError, match='shapes should be equal'):  # branching.py:71-75
    mism[2](x, None, 'nngp')


def test_fused_vs_per_layer_path_full_size(nt):
  """The fused kernels and the one-kernel-per-layer path are independent implementations;
  at BASELINE's Myrtle-10 / 32x32x3 size they must agree (fp32 5e-5, fp64 1e-11), incl. ragged
  block shapes that exercise partial CTA groups and tile edges."""
  x1 = np.random.default_rng(51).standard_normal((7, 32, 32, 3)).astype(np.float32)
  x2 = np.random.default_rng(52).standard_normal((5, 32, 32, 3)).astype(np.float32)
  _, _, kernel_fn = cases.build(cases.myrtle(10), nt.stax)
  for x64, tol in ((False, 5e-5), (True, 1e-11)):
    nt.config.update('enable_x64', x64)
    nt.config.update('disable_fusion', False)
    a = kernel_fn(x1, x2, ('nngp', 'ntk'))
    nt.config.update('disable_fusion', True)
    b = kernel_fn(x1, x2, ('nngp', 'ntk'))
    nt.config.update('disable_fusion', False)
    np.testing.assert_allclose(a.nngp, b.nngp, rtol=tol)
    np.testing.assert_allclose(a.ntk, b.ntk, rtol=tol)


def test_gram_properties_at_scale(nt):
  """Size-independent properties on a 96x96 Myrtle-10 block (fp32): symmetry of K(x, x),
  positive semi-definiteness, block consistency (a sub-block equals the same entries of the
  large block), and exact agreement of duplicated columns."""
  nt.config.update('enable_x64', False)
  _, _, kernel_fn = cases.build(cases.myrtle(10), nt.stax)
  x = np.random.default_rng(61).standard_normal((96, 32, 32, 3)).astype(np.float32)
  k = kernel_fn(x, None, ('nngp', 'ntk'))
  for m in (k.nngp, k.ntk):
    np.testing.assert_allclose(m, m.T, rtol=2e-5)
    w = np.linalg.eigvalsh((m.astype(np.float64) + m.T.astype(np.float64)) / 2)
    assert w.min() > -1e-4 * w.max()
  sub = kernel_fn(x[10:17], x[40:45], ('nngp', 'ntk'))
  np.testing.assert_allclose(sub.nngp, k.nngp[10:17, 40:45], rtol=2e-5)
  np.testing.assert_allclose(sub.ntk, k.ntk[10:17, 40:45], rtol=2e-5)
  x2 = np.concatenate([x[:4], x[:4]])        # duplicated columns give bit-identical entries
  d = kernel_fn(x[20:23], x2, ('nngp', 'ntk'))
  np.testing.assert_array_equal(d.nngp[:, :4], d.nngp[:, 4:])
  np.testing.assert_array_equal(d.ntk[:, :4], d.ntk[:, 4:])


def test_edge_shapes_and_fallbacks(nt):
  """n = 1, non-square images / C != 3 / 28x28 (per-layer fallback), nngp-only fused call."""
  from oracle import ntk_oracle as O
  nt.config.update('enable_x64', True)
  spec = ('serial', [cases.conv(), cases.RELU, cases.conv(), cases.RELU, ('gap',), ('dense', 1.3, 0.1)])
  _, _, kernel_fn = cases.build(spec, nt.stax)
  for shape1, shape2 in (((1, 8, 8, 3), (1, 8, 8, 3)), ((2, 6, 10, 3), (3, 6, 10, 3)),
                         ((2, 8, 8, 1), (1, 8, 8, 1)), ((1, 28, 28, 1), (2, 28, 28, 1)),
                         ((3, 16, 16, 3), (2, 16, 16, 3))):
    x1 = np.random.default_rng(71).standard_normal(shape1).astype(np.float32)
    x2 = np.random.default_rng(72).standard_normal(shape2).astype(np.float32)
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    out = kernel_fn(x1, x2, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, ref[0], rtol=1e-10)
    np.testing.assert_allclose(out.ntk, ref[1], rtol=1e-10)
    np.testing.assert_allclose(kernel_fn(x1, x2, 'nngp'), ref[0], rtol=1e-10)
  nt.config.update('enable_x64', False)


def test_deep_stack_chunking(nt):
  """7 Conv+Relu layers at one resolution are cut into fused chunks 3+3+1 with STORE/LOAD
  boundaries in the sheared layout (README.md:370-386 style net)."""
  from oracle import ntk_oracle as O
  spec = ('serial', [cases.conv(W=1.4, b=0.1), cases.RELU] * 7 + [('gap',), ('dense', 1., 0.)])
  _, _, kernel_fn = cases.build(spec, nt.stax)
  x1 = np.random.default_rng(81).standard_normal((3, 16, 16, 3)).astype(np.float32)
  x2 = np.random.default_rng(82).standard_normal((2, 16, 16, 3)).astype(np.float32)
  ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
  for x64 in (False, True):
    nt.config.update('enable_x64', x64)
    out = kernel_fn(x1, x2, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
    np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
  nt.config.update('enable_x64', False)


def test_batch_multi_device_threads(nt):
  """nt.batch(device_count=D): one host thread + one context per GPU (skips on 1 GPU)."""
  from neural_tangents_b200 import _lib
  if _lib.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  nt.config.update('enable_x64', False)
  _, _, kernel_fn = cases.build(cases.myrtle(5), nt.stax)
  x1 = np.random.default_rng(91).standard_normal((8, 32, 32, 3)).astype(np.float32)
  x2 = np.random.default_rng(92).standard_normal((4, 32, 32, 3)).astype(np.float32)
  full = kernel_fn(x1, x2, ('nngp', 'ntk'))
  out = nt.batch(kernel_fn, batch_size=2, device_count=2)(x1, x2, ('nngp', 'ntk'))
  np.testing.assert_array_equal(out.nngp, full.nngp)
  np.testing.assert_array_equal(out.ntk, full.ntk)
  with pytest.raises(ValueError, match='too small|must divide'):
    nt.batch(kernel_fn, batch_size=2, device_count=2)(x1[:5], x2, 'nngp')


def test_fcn_config1_tensor_core_gram(nt):
  """BASELINE configs[0]: Dense-Relu x3 FCN, 1000 x 1000, 784-d.  The input Gram runs on the
  tensor cores (tcgen05 kind::tf32 with the 3xTF32 split in fp32, DMMA in fp64); ragged sizes
  (1000 and 784 are not tile multiples) exercise the zero-padded edges."""
  from oracle import ntk_oracle as O
  spec = cases.fcn(3)
  x1 = np.random.default_rng(101).standard_normal((1000, 784)).astype(np.float32)
  x2 = np.random.default_rng(102).standard_normal((1000, 784)).astype(np.float32)
  ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
  _, _, kernel_fn = cases.build(spec, nt.stax)
  for x64 in (False, True):
    nt.config.update('enable_x64', x64)
    out = kernel_fn(x1, x2, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
    np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
  # the split must recover (almost) full fp32 accuracy of the raw Gram: relative to the operand
  # scale |x||y|/d ~ 1 a single TF32 pass is off by ~3e-4, 3xTF32 stays below 2e-6
  nt.config.update('enable_x64', False)
  _, _, dense = nt.stax.serial(nt.stax.Dense(1, 1., None))
  g = dense(x1[:300], x2[:200], 'nngp')
  g64 = x1[:300].astype(np.float64) @ x2[:200].astype(np.float64).T / 784
  assert np.abs(g - g64).max() < 2e-6


def test_triangular_schedule_equals_full_square(nt):
  """x2=None computes the upper triangle and mirrors it; NTK_FLAG_FULL_SQUARE computes all
  n x n entries like the reference (batching.py:370).  Both must agree (and be symmetric)."""
  nt.config.update('enable_x64', False)
  _, _, kernel_fn = cases.build(cases.myrtle(7), nt.stax)
  x = np.random.default_rng(111).standard_normal((9, 32, 32, 3)).astype(np.float32)
  tri = kernel_fn(x, None, ('nngp', 'ntk'))
  nt.config.update('full_square', True)
  full = kernel_fn(x, None, ('nngp', 'ntk'))
  nt.config.update('full_square', False)
  np.testing.assert_array_equal(tri.nngp, tri.nngp.T)
  np.testing.assert_array_equal(tri.ntk, tri.ntk.T)
  np.testing.assert_allclose(tri.nngp, full.nngp, rtol=1e-5)
  np.testing.assert_allclose(tri.ntk, full.ntk, rtol=1e-5)
  iu = np.triu_indices(9)
  np.testing.assert_array_equal(tri.nngp[iu], full.nngp[iu])


@pytest.mark.parametrize('size', [16, 32])
def test_wide_resnet_fused_column_sparse_path(nt, size):
  """WideResNet-style nets (README.md:192-222: pre-activation blocks, identity and conv shortcuts,
  stride-2 blocks, FanInSum) go through the column-sparse residual kernels (res_kernels.cuh).
  Checked against the oracle and against the general per-op path."""
  from oracle import ntk_oracle as O
  spec = cases.wrn() if size == 16 else ('serial', [
      cases.conv(W=1., b=0.1), cases.wrn_block(1, True), cases.wrn_block(1, False),
      cases.wrn_block(2, True), cases.wrn_block(1, False), cases.wrn_block(2, True),
      cases.wrn_block(1, False), pool((8, 8), (1, 1)), ('flatten',), ('dense', 1., 0.)])
  x1 = np.random.default_rng(121).standard_normal((3, size, size, 3)).astype(np.float32)
  x2 = np.random.default_rng(122).standard_normal((2, size, size, 3)).astype(np.float32)
  ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
  _, _, kernel_fn = cases.build(spec, nt.stax)
  for x64 in (False, True):
    nt.config.update('enable_x64', x64)
    out = kernel_fn(x1, x2, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
    np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
    nt.config.update('disable_fusion', True)
    gen = kernel_fn(x1, x2, ('nngp', 'ntk'))
    nt.config.update('disable_fusion', False)
    np.testing.assert_allclose(out.nngp, gen.nngp, rtol=RTOL[x64] / 2)
    np.testing.assert_allclose(out.ntk, gen.ntk, rtol=RTOL[x64] / 2)
  nt.config.update('enable_x64', False)
  sym = kernel_fn(x1, None, ('nngp', 'ntk'))
  refs = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
  off = ~np.eye(3, dtype=bool)
  np.testing.assert_allclose(sym.nngp[off], refs[0][off], rtol=1e-4)
  np.testing.assert_allclose(sym.ntk[off], refs[1][off], rtol=1e-4)
  np.testing.assert_allclose(np.diag(sym.ntk), np.diag(refs[1]), rtol=2e-3)


def test_pool_free_flatten_nets_diagonal_path(nt):
  """Pool-free networks ending in Flatten only need the diagonal column (the reference's
  `diagonal_spatial` fast path, linear.py:3381-3437, README.md:399-416): plain, strided and
  residual variants against the oracle and the general path."""
  from oracle import ntk_oracle as O
  specs = {
      'plain': ('serial', [cases.conv(W=1.3, b=0.1), cases.RELU] * 5 + [('flatten',), ('dense', 1.1, 0.2)]),
      'strided': ('serial', [cases.conv(W=1.3, b=0.1), cases.RELU, cases.conv(s=(2, 2)), ('abrelu', 0.1, 1., False),
                             cases.conv(), cases.RELU, ('flatten',), ('dense', 1., 0.)]),
      'residual': ('serial', [cases.conv(W=1., b=0.1), cases.wrn_block(1, True), cases.wrn_block(2, True),
                              cases.wrn_block(1, False), ('flatten',), ('dense', 1., 0.1)]),
  }
  for name, spec in specs.items():
    for size in (8, 32):
      x1 = np.random.default_rng(131).standard_normal((3, size, size, 3)).astype(np.float32)
      x2 = np.random.default_rng(132).standard_normal((4, size, size, 3)).astype(np.float32)
      ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
      _, _, kernel_fn = cases.build(spec, nt.stax)
      for x64 in (False, True):
        nt.config.update('enable_x64', x64)
        out = kernel_fn(x1, x2, ('nngp', 'ntk'))
        np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64], err_msg=f'{name} {size}')
        np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64], err_msg=f'{name} {size}')
        np.testing.assert_allclose(kernel_fn(x1, x2, 'nngp'), ref[0], rtol=RTOL[x64])
      nt.config.update('enable_x64', False)
      sym = kernel_fn(x1, None, ('nngp', 'ntk'))
      np.testing.assert_array_equal(sym.ntk, sym.ntk.T)
      np.testing.assert_allclose(sym.nngp, O.kernel_fn(spec, x1, None, ('nngp',))[0], rtol=1e-4)


ERF = ('erf', 1., 1., 0.)
ERF2 = ('erf', 0.8, 1.3, 0.2)


def test_erf_fused_stage_path(nt):
  """Erf closed form (elementwise.py:67-112) inside the fused stage kernels: Myrtle-style stacks
  with Erf (and mixed Erf / ABRelu) activations at 32x32 and 16x16, against the oracle and the
  general per-op path; fp32 1e-4 / fp64 1e-10."""
  from oracle import ntk_oracle as O
  specs = {
      'erf_myrtle': ('serial', [cases.conv(W=1.2, b=0.1), ERF, cases.conv(W=1.1, b=0.2), ERF2, pool(),
                                cases.conv(), ERF, pool(), cases.conv(W=1.3, b=0.), ERF2, cases.conv(), ERF,
                                ('gap',), ('dense', 1.1, 0.2)]),
      'mixed': ('serial', [cases.conv(W=1.2, b=0.1), ERF2, cases.conv(), cases.RELU, cases.conv(), ERF, pool(),
                           cases.conv(), cases.RELU, cases.conv(W=0.9, b=0.3), ERF2, ('gap',), ('dense', 1., 0.)]),
  }
  for name, spec in specs.items():
    for size in (32, 16):
      x1 = np.random.default_rng(141).standard_normal((3, size, size, 3)).astype(np.float32)
      x2 = np.random.default_rng(142).standard_normal((2, size, size, 3)).astype(np.float32)
      ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
      _, _, kernel_fn = cases.build(spec, nt.stax)
      for x64 in (False, True):
        nt.config.update('enable_x64', x64)
        out = kernel_fn(x1, x2, ('nngp', 'ntk'))
        np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64], err_msg=f'{name} {size} x64={x64}')
        np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64], err_msg=f'{name} {size} x64={x64}')
        np.testing.assert_allclose(kernel_fn(x1, x2, 'nngp'), ref[0], rtol=RTOL[x64])
        nt.config.update('disable_fusion', True)
        gen = kernel_fn(x1, x2, ('nngp', 'ntk'))
        nt.config.update('disable_fusion', False)
        np.testing.assert_allclose(out.ntk, gen.ntk, rtol=RTOL[x64])
      nt.config.update('enable_x64', False)
      sym = kernel_fn(x1, None, ('nngp', 'ntk'))
      refs = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
      # Erf has no sqrt singularity at duplicates: the diagonal holds to the plain tolerance
      np.testing.assert_allclose(sym.nngp, refs[0], rtol=1e-4)
      np.testing.assert_allclose(sym.ntk, refs[1], rtol=1e-4)


@pytest.mark.parametrize('size', [16, 32])
def test_erf_wide_resnet_and_diagonal_paths(nt, size):
  """WideResNet with Erf (BASELINE config 5, Erf variant) through the column-sparse residual kernels,
  and a pool-free Erf net ending in Flatten through the diagonal-column kernel."""
  from oracle import ntk_oracle as O
  wrn = ('serial', [cases.conv(W=1., b=0.1), cases.wrn_block(1, True, ERF), cases.wrn_block(1, False, ERF2),
                    cases.wrn_block(2, True, ERF), cases.wrn_block(1, False, cases.RELU), ('gap',),
                    ('dense', 1., 0.1)])
  flat = ('serial', [cases.conv(W=1.3, b=0.1), ERF, cases.conv(s=(2, 2)), cases.wrn_block(1, False, ERF),
                     ERF2, cases.conv(), cases.RELU, ('flatten',), ('dense', 1.1, 0.2)])
  x1 = np.random.default_rng(151).standard_normal((3, size, size, 3)).astype(np.float32)
  x2 = np.random.default_rng(152).standard_normal((2, size, size, 3)).astype(np.float32)
  for name, spec in (('wrn', wrn), ('flat', flat)):
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    refs = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
    _, _, kernel_fn = cases.build(spec, nt.stax)
    for x64 in (False, True):
      nt.config.update('enable_x64', x64)
      out = kernel_fn(x1, x2, ('nngp', 'ntk'))
      np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64], err_msg=f'{name} x64={x64}')
      np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64], err_msg=f'{name} x64={x64}')
      sym = kernel_fn(x1, None, ('nngp', 'ntk'))
      off = ~np.eye(3, dtype=bool)
      np.testing.assert_allclose(sym.nngp[off], refs[0][off], rtol=RTOL[x64])
      np.testing.assert_allclose(sym.ntk[off], refs[1][off], rtol=RTOL[x64])
      np.testing.assert_allclose(np.diag(sym.ntk), np.diag(refs[1]), rtol=RTOL_DUP[x64])
    nt.config.update('enable_x64', False)


def test_duplicate_pairs_reproduce_self_diagonal(nt):
  """Exact-duplicate pairs (the diagonal of K(x, x), or x2 holding copies of x1 rows) hit the
  sqrt singularity of the arccos kernel: q1 q2 - K^2 must be exactly 0 on their diagonal, which needs
  the self-pair (q-map) pipeline and the cross-pair kernels to agree bit for bit across pooled stage
  boundaries (same arithmetic, same accumulation order).  With that the duplicates meet the ordinary
  tolerance instead of the sqrt(eps) bound of SURVEY Appendix A."""
  from oracle import ntk_oracle as O
  x = np.random.default_rng(161).standard_normal((4, 32, 32, 3)).astype(np.float32)
  d = np.eye(4, dtype=bool)
  for spec in (cases.myrtle(10), cases.myrtle(5, 'gap')):
    ref = O.kernel_fn(spec, x, None, ('nngp', 'ntk'))
    _, _, kernel_fn = cases.build(spec, nt.stax)
    for x64, tol in ((False, 1e-5), (True, 1e-11)):
      nt.config.update('enable_x64', x64)
      sym = kernel_fn(x, None, ('nngp', 'ntk'))
      dup = kernel_fn(x, x.copy(), ('nngp', 'ntk'))
      for out in (sym, dup):
        np.testing.assert_allclose(out.nngp[d], ref[0][d], rtol=tol)
        np.testing.assert_allclose(out.ntk[d], ref[1][d], rtol=tol)
      np.testing.assert_array_equal(np.diag(sym.ntk), np.diag(dup.ntk))
  nt.config.update('enable_x64', False)


def test_packed_and_scalar_stage_kernels_agree(nt, tmp_path):
  """`k_stage_p` (packed FFMA2, U = T + K carry, MUFU.RSQ normalisation) and the scalar `k_stage` are two
  implementations of the same stage; NTK_B200_NO_PACKED=1 (read at library load) selects the scalar one in a
  child process.  They agree far inside the fp32 budget on Myrtle-10 and on a 16x16 stack."""
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  code = (
      "import sys, numpy as np\n"
      f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests', 'golden')!r})\n"
      "import cases, neural_tangents_b200 as nt\n"
      "x1 = np.random.default_rng(171).standard_normal((5, 32, 32, 3)).astype(np.float32)\n"
      "x2 = np.random.default_rng(172).standard_normal((4, 32, 32, 3)).astype(np.float32)\n"
      "_, _, k = cases.build(cases.myrtle(10), nt.stax)\n"
      "a = k(x1, x2, ('nngp', 'ntk')); s = k(x1, None, ('nngp', 'ntk'))\n"
      "_, _, k16 = cases.build(cases.CASES['myrtle10_16px'][0], nt.stax)\n"
      "b = k16(x1[:, ::2, ::2], x2[:, ::2, ::2], ('nngp', 'ntk'))\n"
      "erf = ('serial', [('erf', 1., 1., 0.) if l == cases.RELU else l for l in cases.myrtle(10)[1]])\n"
      "_, _, ke = cases.build(erf, nt.stax)\n"
      "e = ke(x1, x2, ('nngp', 'ntk'))\n"
      "np.savez(sys.argv[1], a0=a.nngp, a1=a.ntk, s0=s.nngp, s1=s.ntk, b0=b.nngp, b1=b.ntk, e0=e.nngp, e1=e.ntk)\n")
  outs = {}
  for tag, env_extra in (('packed', {}), ('scalar', {'NTK_B200_NO_PACKED': '1'})):
    path = str(tmp_path / f'{tag}.npz')
    env = dict(os.environ, **env_extra)
    env.pop('NTK_B200_NO_PACKED', None) if tag == 'packed' else None
    subprocess.run([sys.executable, '-c', code, path], check=True, env=env, timeout=600)
    outs[tag] = np.load(path)
  for key in ('a0', 'a1', 's0', 's1', 'b0', 'b1', 'e0', 'e1'):
    # the cross entries of the bias-free Erf net are ~1e-5 of its diagonal scale (odd activation): compare
    # them on the scale of the kernel (entries of K(x, x) are O(0.1)), not entry by entry
    atol = 0. if key[0] != 'e' else 1e-7
    np.testing.assert_allclose(outs['packed'][key], outs['scalar'][key], rtol=3e-6 if key[0] != 'e' else 1e-5,
                               atol=atol, err_msg=key)


def test_predict_on_gpu_grams(nt):
  """`nt.predict.gradient_descent_mse_ensemble` on Gram matrices produced by the CUDA path (Myrtle-5, batched):
  interpolation of the training targets, agreement of t = None with a direct solve on the oracle's float64
  kernels, positive semi-definite posterior covariance."""
  from oracle import ntk_oracle as O
  spec = cases.myrtle(5, 'gap')
  _, _, kernel_fn = cases.build(spec, nt.stax)
  rng = np.random.default_rng(181)
  x_train = rng.standard_normal((12, 32, 32, 3)).astype(np.float32)
  x_test = rng.standard_normal((4, 32, 32, 3)).astype(np.float32)
  y_train = rng.standard_normal((12, 10))
  nt.config.update('enable_x64', True)
  batched = nt.batch(kernel_fn, batch_size=4, device_count=0)
  fn = nt.predict.gradient_descent_mse_ensemble(batched, x_train, y_train, diag_reg=1e-6)
  out = fn(t=None, x_test=x_test, get=('nngp', 'ntk'), compute_cov=True)
  k_dd = O.kernel_fn(spec, x_train, None, ('nngp', 'ntk'))
  k_td = O.kernel_fn(spec, x_test, x_train, ('nngp', 'ntk'))
  for i, g in enumerate(('nngp', 'ntk')):
    A = k_dd[i] + 1e-6 * np.trace(k_dd[i]) / 12 * np.eye(12)
    np.testing.assert_allclose(getattr(out, g).mean, k_td[i] @ np.linalg.solve(A, y_train), rtol=1e-6, atol=1e-9)
    c = getattr(out, g).covariance
    assert np.linalg.eigvalsh((c + c.T) / 2).min() > -1e-9
  np.testing.assert_allclose(fn(t=None, x_test=x_train, get='ntk'), y_train, rtol=1e-2, atol=2e-3)  # regularised fit
  near = fn(t=1e10, x_test=x_test, get='ntk')
  np.testing.assert_allclose(near, out.ntk.mean, rtol=1e-5, atol=1e-8)
  nt.config.update('enable_x64', False)
