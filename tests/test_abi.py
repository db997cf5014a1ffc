"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/ntk_b200.h declares, and the host-side program logic works without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'ntk_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(ntk_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
  import __graft_entry__ as g
  g.build()
  from neural_tangents_b200 import _lib
  return _lib


def test_library_exports_every_declared_symbol(lib):
  cdll = lib.load()
  declared = _declared_symbols()
  assert len(declared) >= 15
  for name in declared:
    assert hasattr(cdll, name), f'{name} declared in include/ntk_b200.h but not exported'
  assert set(declared) == set(lib.EXPORTED_SYMBOLS)
  assert cdll.ntk_abi_version() == 2


def test_struct_layout_matches_header(lib):
  # ntk_op_t: 4 + 6 int32 (40 bytes) + 4 doubles = 72 bytes; ntk_state_t: 4 ptrs + 6 int32 = 56
  assert ctypes.sizeof(lib.NtkOp) == 72
  assert ctypes.sizeof(lib.NtkState) == 56


def test_program_shape_inference_without_gpu(lib):
  import cases
  from neural_tangents_b200 import stax
  _, _, kf = cases.build(cases.myrtle(10), stax)
  low = stax._lowered(stax._strip(kf._spec), False, False, True)
  assert low.program.output_shape(32, 32) == (0, 0, True)
  _, _, kf = cases.build(('serial', [cases.conv(), cases.RELU, cases.pool()]), stax)
  low = stax._lowered(stax._strip(kf._spec), False, False, True)
  assert low.program.output_shape(32, 32) == (16, 16, False)
  assert low.out_meta.is_reversed is True
  # WideResNet: FanOut/parallel/FanInSum lower to slot reuse + NTK_OP_FANINSUM
  _, _, kf = cases.build(cases.wrn(), stax)
  low = stax._lowered(stax._strip(kf._spec), False, False, True)
  assert sum(1 for op in low.ops if op[0] == lib.OP_FANINSUM) == 4
  assert low.program.output_shape(8, 8) == (0, 0, True)


def test_errors_without_gpu(lib):
  from neural_tangents_b200 import stax
  _, _, kf = stax.serial(stax.Relu())
  low = stax._lowered(stax._strip(kf._spec), False, False, False)
  with pytest.raises(ValueError, match='must be Gaussian'):      # elementwise.py:1267-1270
    low.program.output_shape(0, 0)
  with pytest.raises(NotImplementedError):
    stax.Dense(1, parameterization='standard')
  with pytest.raises(NotImplementedError):
    stax.Conv(1, (3, 3), dimension_numbers=('NCHW', 'OIHW', 'NCHW'))
  _, _, kf = stax.serial(stax.Dense(1))
  with pytest.raises(TypeError):                                   # requirements.py:754-758
    kf([[1., 2.]], None, 'nngp')
  with pytest.raises(ValueError, match='unique'):                  # utils.py:150-152
    kf(np.ones((2, 3)), None, ('nngp', 'NNGP'))
  with pytest.raises(ValueError, match='non-empty'):
    kf(np.ones((2, 3)), None, ())
  with pytest.raises(NotImplementedError):
    kf(np.ones((2, 3)), None, 'nngp', mask_constant=0.)


def test_no_gpu_means_loud_failure(lib):
  """The product path has no CPU fallback."""
  n = ctypes.c_int(0)
  st = lib.load().ntk_device_count(ctypes.byref(n))
  if st == 0 and n.value > 0:
    pytest.skip('a GPU is present')
  from neural_tangents_b200 import stax
  _, _, kf = stax.serial(stax.Dense(1), stax.Relu(), stax.Dense(1))
  with pytest.raises(Exception, match='no CUDA device|CUDA'):
    kf(np.ones((2, 3), np.float32), None, 'nngp')


def test_batch_size_arithmetic():
  """batching.py:647-679."""
  from neural_tangents_b200 import batching
  assert batching._get_n_batches_and_batch_sizes(12, 8, 4, 1) == (3, 4, 2, 4)
  assert batching._get_n_batches_and_batch_sizes(16, 8, 2, 4) == (2, 8, 4, 2)
  with pytest.raises(ValueError, match='rows of kernel must divide'):
    batching._get_n_batches_and_batch_sizes(10, 10, 3, 1)
  with pytest.warns(UserWarning, match='Batch size is reduced'):
    assert batching._get_n_batches_and_batch_sizes(8, 8, 16, 1) == (1, 8, 1, 8)
  with pytest.raises(ValueError, match='must divide number of physical devices'):
    batching._get_n_per_device(10, 4)


def test_batch_stitching_with_foreign_kernel_fn():
  """`batch` wraps any kernel_fn(x1, x2, ...) (batching.py:88-92): batched == unbatched."""
  from neural_tangents_b200 import batching

  def kernel_fn(x1, x2=None, get=None):
    x2 = x1 if x2 is None else x2
    return x1 @ x2.T

  x1 = np.random.default_rng(0).standard_normal((12, 5))
  x2 = np.random.default_rng(1).standard_normal((8, 5))
  b = batching.batch(kernel_fn, batch_size=4, device_count=0)
  np.testing.assert_allclose(b(x1, x2), kernel_fn(x1, x2), rtol=1e-13, atol=1e-14)
  np.testing.assert_allclose(b(x1), kernel_fn(x1), rtol=1e-13, atol=1e-14)


def test_bench_reference_arm_prints_one_json_line():
  """`bench.py --impl reference` (the CPU arm of the driver contract) prints exactly one JSON line on stdout with
  the contract's keys; everything else (including library banners) goes to stderr."""
  import json
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--workload', 'fcn',
                        '--steps', '1', '--warmup', '0', '--ref-cols', '1'], capture_output=True, text=True,
                       timeout=300, check=True).stdout
  lines = [l for l in out.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
              'scaling', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
    assert key in d, key
  assert d['impl'] == 'reference' and d['cpu_baseline']['kind'] == 'port' and d['value'] > 0
  assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0


def test_kernel_family_selection_without_gpu(lib):
  """`ntk_program_path`: which kernel family a network runs on (host logic of the planners).  Guards against a
  BASELINE configuration silently dropping to the one-kernel-per-op path."""
  import cases
  from neural_tangents_b200 import stax

  def path(spec, H, W, C, **kw):
    _, _, kf = cases.build(spec, stax)
    low = stax._lowered(stax._strip(kf._spec), False, False, H > 0)
    return low.program.path(H, W, C, **kw)

  erf = ('erf', 1., 1., 0.)
  erf_myrtle = ('serial', [erf if l == cases.RELU else l for l in cases.myrtle(10)[1]])
  for x64 in (False, True):
    for depth in (5, 7, 10):                                                   # BASELINE configs 2-4
      assert path(cases.myrtle(depth), 32, 32, 3, x64=x64) == 'fused'
      assert path(cases.myrtle(depth, 'gap'), 32, 32, 3, x64=x64) == 'fused'
    assert path(erf_myrtle, 32, 32, 3, x64=x64) == 'fused'                      # Erf closed form is fused too
    two_stage_16 = ('serial', [cases.conv(), cases.RELU] * 3 + [cases.pool()] + [cases.conv(), cases.RELU] * 3 +
                    [('gap',), ('dense', 1., 0.)])
    assert path(two_stage_16, 16, 16, 3, x64=x64) == 'fused'                     # 16 -> 8
    assert path(cases.CASES['myrtle10_16px'][0], 16, 16, 3, x64=x64) == 'generic'  # 16 -> 8 -> 4: no 4x4 stage kernel
    assert path(cases.fcn(3), 0, 0, 784, x64=x64) == 'fcn'                      # configs[0]
    assert path(cases.wrn(), 16, 16, 3, x64=x64) == 'res'                       # config 5 (Relu and Erf)
    assert path(cases.wrn(erf), 32, 32, 3, x64=x64) == 'res'
  flat = ('serial', [cases.conv(W=1., b=None), cases.RELU] * 21 + [('flatten',)])   # README.md:399-416
  assert path(flat, 32, 32, 3) == 'diag'
  assert path(('serial', [cases.conv(W=1., b=None), cases.RELU] * 21 + [('gap',)]), 32, 32, 3) == 'fused'
  # outside the fused families: other image sizes / channel counts, VALID convs, forced per-op path, cov outputs
  assert path(cases.myrtle(10), 28, 28, 1) == 'generic'
  assert path(cases.CASES['erf_valid'][0], 6, 5, 2) == 'generic'
  assert path(cases.myrtle(10), 32, 32, 3, flags=lib.FLAG_NO_FUSION) == 'generic'
  assert path(cases.myrtle(10), 32, 32, 3, flags=lib.FLAG_WANT_COV) == 'generic'
  assert path(cases.CASES['wrn_relu'][0], 8, 8, 3) == 'generic'                 # 8x8 with a stride-2 block: 4x4 maps


def test_comm_entry_points_without_gpu(lib):
  """libnccl is resolved with dlopen on first use (no link-time dependency); without a GPU the communicator
  cannot be created, but the version query and argument checks work."""
  cdll = lib.load()
  v = ctypes.c_int()
  rc = cdll.ntk_comm_nccl_version(ctypes.byref(v))
  if rc == 0:
    assert v.value >= 21800
  else:
    assert rc == lib.E_UNSUPPORTED and b'NCCL' in cdll.ntk_last_error()
  out = ctypes.c_void_p()
  assert cdll.ntk_comm_create(None, None, 0, 1, ctypes.byref(out)) == lib.E_INVAL
  assert cdll.ntk_comm_rank(None) == -1 and cdll.ntk_comm_world(None) == -1
  assert cdll.ntk_context_device(None) == -1
  assert cdll.ntk_gram_device_on_stream(None, None, 0, None, 1, None, 1, 0, 0, 1, 0, None, None, 1, None, None,
                                        None) == lib.E_INVAL
  assert cdll.ntk_apply_device(None, None, 0, None, None, None) == lib.E_INVAL
  assert cdll.ntk_sym_assemble(None, 0, None, 0, None, 0, None, 0) == lib.E_INVAL
  # pinned allocations fall back to pageable NumPy memory when there is no CUDA device
  a = lib.pinned_empty((512, 512), np.float32)
  assert a.shape == (512, 512) and a.flags['C_CONTIGUOUS'] and a.flags['WRITEABLE']


def test_diag_path_selection_covers_odd_sizes(lib):
  """Pool-free Flatten nets take the diagonal-column path at any square size whose images fit shared memory
  (odd sizes included: SAME stride-2 geometry is handled in the kernels), else the per-op path."""
  import cases
  from neural_tangents_b200 import stax
  spec = ('serial', [cases.conv(), cases.RELU, cases.conv(s=(2, 2)), cases.RELU, ('flatten',), ('dense', 1., 0.)])
  _, _, kf = cases.build(spec, stax)
  low = stax._lowered(stax._strip(kf._spec), False, False, True)
  for size in (7, 15, 28, 32, 64):
    assert low.program.path(size, size, 1) == 'diag'
  assert low.program.path(64, 64, 1, x64=True) == 'generic'     # 8 x 64 x 64 doubles do not fit
  assert low.program.path(96, 96, 1) == 'generic'
  assert low.program.path(28, 14, 1) == 'generic'


def test_fused_path_covers_grey_nonsquare_and_mnist_sizes(lib):
  """Round 2: the fused stage kernels take any H x W <= 32 x 32 with 1 or 3 channels (embedded in the next
  shear size); what the embedding cannot express (SAME pools on odd sizes, > 32 px, other channel counts, Erf on
  embedded sizes) stays on the per-op path."""
  import cases
  from neural_tangents_b200 import stax

  def path(spec, H, W, C, **kw):
    _, _, kf = cases.build(spec, stax)
    low = stax._lowered(stax._strip(kf._spec), False, False, True)
    return low.program.path(H, W, C, **kw)
  gap5 = cases.myrtle(5, 'gap')
  assert path(gap5, 28, 28, 1) == 'fused'                     # MNIST
  assert path(gap5, 28, 28, 1, x64=True) == 'fused'
  assert path(cases.myrtle(5), 32, 32, 1) == 'fused'          # grey CIFAR-sized
  assert path(cases.CASES['conv_pool_20x12'][0], 20, 12, 3) == 'fused'
  assert path(cases.CASES['conv_pool_15_odd'][0], 15, 15, 1) == 'fused'      # VALID pool on an odd size: floor
  assert path(gap5, 24, 24, 3) == 'fused'
  assert path(gap5, 28, 28, 2) == 'fused'                     # other channel counts: k_input_shear + LOADing first stage
  assert path(gap5, 32, 32, 16) == 'fused' and path(cases.myrtle(10), 32, 32, 4, x64=True) == 'fused'
  assert path(gap5, 32, 32, 128) == 'fused' and path(gap5, 32, 32, 512) == 'generic'   # pre-pass rows must fit shared memory
  assert path(gap5, 40, 40, 3) == 'generic'                   # larger than the largest shear
  same_pool = ('serial', [cases.conv(), cases.RELU, cases.pool(pad='SAME'), cases.conv(), cases.RELU, ('gap',)])
  assert path(same_pool, 16, 16, 3) == 'fused'                # SAME == VALID on even sizes
  assert path(same_pool, 15, 15, 3) == 'generic'              # SAME pads an odd size with zeros
  erf5 = ('serial', [('erf', 1., 1., 0.) if l == cases.RELU else l for l in gap5[1]])
  assert path(erf5, 32, 32, 3) == 'fused' and path(erf5, 28, 28, 1) == 'fused'      # EMB general-activation family
  assert path(erf5, 32, 32, 1) == 'fused' and path(erf5, 32, 32, 5) == 'fused'   # shear sizes, any C: pre-pass + Erf family
  # 3x3 / 1 / VALID convs: Flatten nets on the diagonal-column kernels, pooled / GAP nets on the stage kernels when every pool sits
  # behind an even number of VALID convs (the box origin must be even), otherwise per op
  V = lambda: cases.conv(pad='VALID')
  assert path(('serial', [V(), cases.RELU] * 3 + [('flatten',), ('dense', 1., 0.)]), 12, 12, 3) == 'diag'
  assert path(cases.CASES['circular_flatten'][0], 9, 9, 2) == 'diag'                                        # CIRCULAR: taps mod S
  assert path(('serial', [V(), cases.RELU] * 2 + [cases.pool(), V(), cases.RELU, ('gap',)]), 28, 28, 1) == 'fused'
  assert path(('serial', [V(), cases.RELU] * 3 + [cases.pool(), V(), cases.RELU, ('gap',)]), 16, 16, 3) == 'generic'
  assert path(('serial', [V(), cases.RELU, cases.conv(), cases.RELU, ('gap',)]), 16, 16, 3) == 'generic'   # mixed paddings
  assert path(('serial', [V(), cases.RELU] * 4 + [('gap',)]), 8, 8, 3) == 'generic'                         # 8 - 2 * 4: empty map
  # SumPool / GlobalSumPool ride the same kernels (epilogue scales)
  sp = ('serial', [cases.conv(), cases.RELU, ('sumpool', (2, 2), (2, 2), 'VALID'), cases.conv(), cases.RELU, ('gsp',)])
  assert path(sp, 32, 32, 3) == 'fused' and path(sp, 28, 28, 1) == 'fused'
  # Gelu / Sin / Rbf stages: the general family of the fused stage kernels at the native sizes, per-op elsewhere
  assert path(cases.CASES['gelu_conv'][0], 32, 32, 3) == 'fused'
  assert path(cases.CASES['rbf_conv_pool'][0], 16, 16, 3) == 'fused'
  assert path(cases.CASES['gelu_conv'][0], 28, 28, 1) == 'fused'
  assert path(cases.CASES['layernorm_conv'][0], 32, 32, 3) == 'generic'
