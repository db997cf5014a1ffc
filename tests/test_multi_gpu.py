"""GPU tests of the round-2 boundary: the NCCL communicator inside libntk_b200.so and the torch-free
`distributed.gram` (contiguous rows for x2 given, folded-cyclic triangular schedule for x2=None), the
NTK_FLAG_UPPER_ONLY trapezoid, the caller-stream entry, device-side Kernel-in/Kernel-out, pinned host
memory, and the diagonal-column path at odd sizes.  World-size-2 cases need two GPUs and skip otherwise."""
import ctypes
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = {False: 1e-4, True: 1e-10}
RTOL_DUP = {False: 2e-3, True: 2e-7}


@pytest.fixture(scope='module')
def nt():
  import __graft_entry__ as g
  g.build()
  import neural_tangents_b200 as nt
  yield nt
  nt.config.update('enable_x64', False)
  nt.config.update('disable_fusion', False)


def _check_sym(v, ref, x64):
  off = ~np.eye(v.shape[0], dtype=bool)
  np.testing.assert_allclose(v[off], ref[off], rtol=RTOL[x64])
  np.testing.assert_allclose(np.diag(v), np.diag(ref), rtol=RTOL_DUP[x64])
  np.testing.assert_array_equal(v, v.T)


def test_distributed_gram_world1_nccl(nt):
  """The product multi-GPU path on a 1-rank NCCL communicator: device broadcast, per-block
  `ntk_gram_device`, all-gather, `ntk_sym_assemble`, all through the C-ABI, against the oracle."""
  from neural_tangents_b200 import _lib, distributed as D
  from oracle import ntk_oracle as O
  assert _lib.Comm.nccl_version() >= 21800
  spec = cases.myrtle(5)
  _, _, kernel_fn = cases.build(spec, nt.stax)
  x1 = np.random.default_rng(3).standard_normal((7, 32, 32, 3)).astype(np.float32)
  x2 = np.random.default_rng(4).standard_normal((5, 32, 32, 3)).astype(np.float32)
  be = D.init(rank=0, world=1, local_rank=0, unique_id=_lib.Comm.unique_id())
  try:
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
    for x64 in (False, True):
      nt.config.update('enable_x64', x64)
      out = D.gram(kernel_fn, x1, x2, ('nngp', 'ntk'))
      np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
      np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
      for block in (None, 3):                       # one trapezoid / ragged trapezoids 3 + 3 + 1
        sym = D.gram(kernel_fn, x1, None, ('nngp', 'ntk'), block_rows=block)
        _check_sym(sym.nngp, sref[0], x64)
        _check_sym(sym.ntk, sref[1], x64)
      only = D.gram(kernel_fn, x1, x2, 'ntk')
      np.testing.assert_allclose(only, ref[1], rtol=RTOL[x64])
      dev = D.gram(kernel_fn, x1, x2, 'nngp', to_host=False)       # stays in HBM
      assert isinstance(dev, D.DeviceArray) and dev.shape == (7, 5)
      np.testing.assert_allclose(be.download(dev), ref[0], rtol=RTOL[x64])
  finally:
    nt.config.update('enable_x64', False)
    D.shutdown()


_WORLD2 = textwrap.dedent('''
    import os, sys
    import numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests', 'golden'))
    import neural_tangents_b200 as nt
    from neural_tangents_b200 import distributed as D
    from oracle import ntk_oracle as O
    import cases
    rank = int(os.environ['RANK'])
    be = D.init()
    assert be.world == 2 and be.ctx.device == rank
    spec = cases.myrtle(5)
    _, _, kernel_fn = cases.build(spec, nt.stax)
    x1 = np.random.default_rng(3).standard_normal((8, 32, 32, 3)).astype(np.float32)
    x2 = np.random.default_rng(4).standard_normal((5, 32, 32, 3)).astype(np.float32)
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
    out = D.gram(kernel_fn, x1 if rank == 0 else None, x2 if rank == 0 else None, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, ref[0], rtol=1e-4)
    np.testing.assert_allclose(out.ntk, ref[1], rtol=1e-4)
    for block in (None, 1, 3):
      sym = D.gram(kernel_fn, x1 if rank == 0 else None, None, ('nngp', 'ntk'), block_rows=block)
      off = ~np.eye(8, dtype=bool)
      np.testing.assert_allclose(sym.ntk[off], sref[1][off], rtol=1e-4)
      np.testing.assert_allclose(np.diag(sym.ntk), np.diag(sref[1]), rtol=2e-3)
      np.testing.assert_allclose(sym.nngp, sref[0], rtol=1e-4)
    D.shutdown()
    print('RANK_OK', rank)
''')


def test_distributed_gram_world2_nccl(nt, tmp_path):
  """Two ranks, two GPUs, NCCL over NVLink: inputs only on rank 0, results identical to the oracle on both."""
  from neural_tangents_b200 import _lib
  if _lib.device_count() < 2:
    pytest.skip('needs two GPUs')
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  script = tmp_path / 'w2.py'
  script.write_text(_WORLD2.format(root=ROOT))
  procs = []
  for r in range(2):
    env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1',
               MASTER_PORT=str(port), NTK_B200_RENDEZVOUS_DIR=str(tmp_path))
    procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True))
  outs = [p.communicate(timeout=900)[0] for p in procs]
  for r, (p, o) in enumerate(zip(procs, outs)):
    assert p.returncode == 0 and f'RANK_OK {r}' in o, o[-3000:]


def test_upper_only_trapezoid(nt):
  """NTK_FLAG_UPPER_ONLY: rows [r0, r1) of a symmetric Gram against columns [r0, n): entries j >= i match the
  full computation bit for bit, entries j < i are left untouched."""
  from neural_tangents_b200 import _lib, stax
  spec = cases.myrtle(5)
  _, _, kernel_fn = cases.build(spec, stax)
  n, r0, r1 = 9, 2, 6
  x = np.random.default_rng(8).standard_normal((n, 32, 32, 3)).astype(np.float32)
  full = kernel_fn(x, x.copy(), ('nngp', 'ntk'))          # cross-pair path on all n x n entries
  ctx = _lib.get_context()
  low = stax._lowered(stax._strip(kernel_fn._spec), False, False, True)
  dx = ctx.malloc(x.nbytes)
  ctx.h2d(dx, x)
  nb = (r1 - r0) * (n - r0) * 4
  dk, dt_ = ctx.malloc(nb), ctx.malloc(nb)
  ctx.memset(dk, 0xff, nb)                                  # NaN pattern
  ctx.memset(dt_, 0xff, nb)
  row = 32 * 32 * 3 * 4
  _lib.gram_device(ctx, low.program, np.float32, dx + r0 * row, r1 - r0, dx + r0 * row, n - r0, 32, 32, 3,
                   _lib.FLAG_UPPER_ONLY, dk, dt_, n - r0)
  k = ctx.d2h(np.empty((r1 - r0, n - r0), np.float32), dk)
  t = ctx.d2h(np.empty((r1 - r0, n - r0), np.float32), dt_)
  for ptr in (dx, dk, dt_):
    ctx.free(ptr)
  iu = np.triu(np.ones((r1 - r0, n - r0), bool))
  np.testing.assert_array_equal(k[iu], full.nngp[r0:r1, r0:][iu])
  np.testing.assert_array_equal(t[iu], full.ntk[r0:r1, r0:][iu])
  assert np.isnan(k[~iu]).all() and np.isnan(t[~iu]).all()


def test_caller_stream_entry_is_asynchronous(nt):
  """`ntk_gram_device_on_stream` enqueues on a foreign stream and returns before the work has finished; the
  result equals the context-stream result bit for bit.  A following call on the context stream is ordered
  behind it on the device (shared workspace), with no host synchronisation in between."""
  from neural_tangents_b200 import _lib, stax
  lib = _lib.load()
  spec = cases.myrtle(10)
  _, _, kernel_fn = cases.build(spec, stax)
  low = stax._lowered(stax._strip(kernel_fn._spec), False, False, True)
  n1, n2 = 48, 48
  x1 = np.random.default_rng(1).standard_normal((n1, 32, 32, 3)).astype(np.float32)
  x2 = np.random.default_rng(2).standard_normal((n2, 32, 32, 3)).astype(np.float32)
  ref = kernel_fn(x1, x2, ('nngp', 'ntk'))
  ctx = _lib.get_context()
  stream = ctypes.c_void_p()
  _lib.check(lib.ntk_stream_create(ctx.device, ctypes.byref(stream)))
  d1, d2 = ctx.malloc(x1.nbytes), ctx.malloc(x2.nbytes)
  ctx.h2d(d1, x1)
  ctx.h2d(d2, x2)
  ctx.synchronize()
  outs = [ctx.malloc(n1 * n2 * 4) for _ in range(4)]
  _lib.gram_device(ctx, low.program, np.float32, d1, n1, d2, n2, 32, 32, 3, 0, outs[0], outs[1], n2,
                   stream=stream.value)
  done = ctypes.c_int32(-1)
  _lib.check(lib.ntk_stream_query(stream, ctypes.byref(done)))
  assert done.value == 0, 'the call blocked until the GPU work was finished'      # ~10 ms of kernels enqueued
  # second call on the context's own stream: must wait for the first on the device (same workspace)
  _lib.gram_device(ctx, low.program, np.float32, d1, n1, d2, n2, 32, 32, 3, 0, outs[2], outs[3], n2)
  _lib.check(lib.ntk_stream_synchronize(stream))
  ctx.synchronize()
  got = [ctx.d2h(np.empty((n1, n2), np.float32), p) for p in outs]
  np.testing.assert_array_equal(got[0], ref.nngp)
  np.testing.assert_array_equal(got[1], ref.ntk)
  np.testing.assert_array_equal(got[2], ref.nngp)
  np.testing.assert_array_equal(got[3], ref.ntk)
  for p in outs + [d1, d2]:
    ctx.free(p)
  lib.ntk_stream_destroy(stream)


@pytest.mark.parametrize('x64', [False, True])
def test_apply_device_matches_apply_host(nt, x64):
  """Kernel-in / Kernel-out on device pointers (`ntk_apply_device`) == the host-pointer entry; the caller's
  input tensors are not modified (copy-on-write inside the executor)."""
  from neural_tangents_b200 import _lib, stax
  nt.config.update('enable_x64', x64)
  dt = np.float64 if x64 else np.float32
  rng = np.random.default_rng(5)
  x1 = rng.standard_normal((2, 6, 6, 2)).astype(np.float32)
  x2 = rng.standard_normal((3, 6, 6, 2)).astype(np.float32)
  a = stax.serial(stax.Conv(1, (3, 3), padding='SAME', W_std=1.2, b_std=0.1), stax.Relu())
  b = stax.serial(stax.Conv(1, (3, 3), padding='SAME', W_std=1.0, b_std=0.3), stax.Relu(), stax.GlobalAvgPool(),
                  stax.Dense(1, 1.1, 0.2))
  k_mid = a[2](x1, x2)
  want = b[2](k_mid)                                           # host path
  ctx = _lib.get_context()
  rev = bool(k_mid.is_reversed)
  can = lambda m, nb: np.ascontiguousarray(stax._from_ref_layout(np.asarray(m, dt), rev, nb))
  ins = [can(k_mid.nngp, 2), can(k_mid.ntk, 2), can(k_mid.cov1, 1), can(k_mid.cov2, 1)]
  d_in = []
  for arr in ins:
    p = ctx.malloc(arr.nbytes)
    ctx.h2d(p, arr)
    d_in.append(p)
  low = stax._lowered(stax._strip(b[2]._spec), rev, True, True)
  isz = np.dtype(dt).itemsize
  d_out = [ctx.malloc(2 * 3 * isz), ctx.malloc(2 * 3 * isz), ctx.malloc(2 * isz), ctx.malloc(3 * isz)]
  mode, gauss = _lib.apply_device(ctx, low.program, dt, 2, 3, 6, 6, _lib.NTK_TENSOR, k_mid.is_gaussian, d_in, d_out,
                                  0, 0)
  ctx.synchronize()
  assert mode == _lib.NTK_TENSOR and gauss
  got = [ctx.d2h(np.empty(s, dt), p) for s, p in zip([(2, 3), (2, 3), (2,), (3,)], d_out)]
  np.testing.assert_array_equal(got[0], want.nngp)
  np.testing.assert_array_equal(got[1], want.ntk)
  np.testing.assert_array_equal(got[2], want.cov1)
  np.testing.assert_array_equal(got[3], want.cov2)
  for arr, p in zip(ins, d_in):
    np.testing.assert_array_equal(ctx.d2h(np.empty_like(arr), p), arr)   # inputs untouched
  for p in d_in + d_out:
    ctx.free(p)
  nt.config.update('enable_x64', False)


def test_pinned_results_and_pageable_staging(nt):
  """Results come back in page-locked memory (direct DMA); pageable inputs larger than one staging slot go
  through the pinned ring; both give the same numbers as pinned inputs."""
  from neural_tangents_b200 import _lib, stax
  lib = _lib.load()
  _, _, kernel_fn = cases.build(cases.fcn(3, 2., 0.05), stax)
  n, d = 3000, 784                                             # 9.4 MB per input > the 8 MB slot
  x1 = np.random.default_rng(0).standard_normal((n, d)).astype(np.float32)
  x2 = np.random.default_rng(1).standard_normal((n, d)).astype(np.float32)
  out = kernel_fn(x1, x2, ('nngp', 'ntk'))
  p1, p2 = _lib.pinned_copy(x1), _lib.pinned_copy(x2)
  out_p = kernel_fn(p1, p2, ('nngp', 'ntk'))
  np.testing.assert_array_equal(out.nngp, out_p.nngp)
  np.testing.assert_array_equal(out.ntk, out_p.ntk)

  class Attr(ctypes.Structure):
    _fields_ = [('type', ctypes.c_int), ('device', ctypes.c_int), ('dptr', ctypes.c_void_p),
                ('hptr', ctypes.c_void_p)]
  try:
    rt = ctypes.CDLL('libcudart.so.12')
  except OSError:
    rt = None
  for arr, pinned in ((out.nngp, True), (p1, True), (x1, False)):
    if rt is None:
      break
    at = Attr()
    rc = rt.cudaPointerGetAttributes(ctypes.byref(at), ctypes.c_void_p(arr.ctypes.data))
    assert rc == 0
    assert (at.type == 1) == pinned, (at.type, pinned)         # cudaMemoryTypeHost == 1
  from oracle import ntk_oracle as O
  ref = O.kernel_fn(cases.fcn(3, 2., 0.05), x1[:64], x2[:64], ('nngp', 'ntk'))
  np.testing.assert_allclose(out.ntk[:64, :64], ref[1], rtol=1e-4)


def test_diagonal_path_odd_and_mnist_sizes(nt):
  """Pool-free Flatten nets with stride-2 convs at odd resolutions (SAME: out = ceil(S/2), window centred at
  2a) and at 28x28 (28 -> 14 -> 7 -> 4): the diagonal-column kernels against the oracle and the per-op path."""
  from oracle import ntk_oracle as O
  spec = ('serial', [cases.conv(W=1.3, b=0.1), cases.RELU, cases.conv(s=(2, 2), W=1.2, b=0.05), cases.RELU,
                     cases.conv(s=(2, 2)), ('abrelu', 0.1, 1., False), cases.conv(s=(2, 2)), cases.RELU,
                     ('flatten',), ('dense', 1., 0.1)])
  _, _, kernel_fn = cases.build(spec, nt.stax)
  for size, C in ((7, 3), (15, 2), (28, 1), (9, 3)):
    x1 = np.random.default_rng(21).standard_normal((3, size, size, C)).astype(np.float32)
    x2 = np.random.default_rng(22).standard_normal((4, size, size, C)).astype(np.float32)
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
    assert low.program.path(size, size, C) == 'diag'
    for x64 in (False, True):
      nt.config.update('enable_x64', x64)
      out = kernel_fn(x1, x2, ('nngp', 'ntk'))
      np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64], err_msg=f'{size}')
      np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64], err_msg=f'{size}')
      nt.config.update('disable_fusion', True)
      gen = kernel_fn(x1, x2, ('nngp', 'ntk'))
      nt.config.update('disable_fusion', False)
      np.testing.assert_allclose(out.ntk, gen.ntk, rtol=RTOL[x64] * 0.1)
  nt.config.update('enable_x64', False)


def test_batch_device_count_one_context_per_device(nt):
  """`batch(device_count=D)` spawns fresh host threads per call; contexts are cached per device, so repeated
  calls do not allocate new workspaces."""
  from neural_tangents_b200 import _lib
  _, _, kernel_fn = cases.build(cases.myrtle(5), nt.stax)
  x1 = np.random.default_rng(0).standard_normal((4, 32, 32, 3)).astype(np.float32)
  x2 = np.random.default_rng(1).standard_normal((2, 32, 32, 3)).astype(np.float32)
  D = min(2, _lib.device_count())
  bk = nt.batch(kernel_fn, batch_size=2, device_count=D)
  ref = kernel_fn(x1, x2, 'ntk')
  n_ctx = None
  for _ in range(4):
    np.testing.assert_array_equal(bk(x1, x2, 'ntk'), ref)
    if n_ctx is None:
      n_ctx = len(_lib._contexts)
    assert len(_lib._contexts) == n_ctx <= max(D, 1) + 1


def test_gram_to_disk_on_gpu_symmetric_trapezoids(nt, tmp_path):
  """SURVEY §8f row 2 on the GPU: restartable slab-by-slab symmetric Gram; slabs are NTK_FLAG_UPPER_ONLY
  trapezoids; the assembled matrices equal the one-call result and the oracle."""
  from neural_tangents_b200.batching import gram_to_disk
  from oracle import ntk_oracle as O
  spec = cases.myrtle(5)
  _, _, kernel_fn = cases.build(spec, nt.stax)
  x = np.random.default_rng(12).standard_normal((7, 32, 32, 3)).astype(np.float32)
  out = gram_to_disk(kernel_fn, x, None, ('nngp', 'ntk'), str(tmp_path), block_rows=3)
  one = kernel_fn(x, None, ('nngp', 'ntk'))
  np.testing.assert_array_equal(np.asarray(out.nngp), one.nngp)
  np.testing.assert_array_equal(np.asarray(out.ntk), one.ntk)
  _check_sym(np.asarray(out.ntk), O.kernel_fn(spec, x, None, ('nngp', 'ntk'))[1], False)
  again = gram_to_disk(kernel_fn, x, None, ('nngp', 'ntk'), str(tmp_path), block_rows=3)   # nothing recomputed
  np.testing.assert_array_equal(np.asarray(again.ntk), one.ntk)
  with pytest.raises(ValueError, match='x1_sha1'):
    gram_to_disk(kernel_fn, x + 1, None, ('nngp', 'ntk'), str(tmp_path), block_rows=3)


def test_batch_symmetric_over_two_devices_is_triangular(nt):
  """`batch(device_count=2)(x, None)`: the triangular folded-cyclic schedule inside one process."""
  from neural_tangents_b200 import _lib
  if _lib.device_count() < 2:
    pytest.skip('needs two GPUs')
  spec = cases.myrtle(5)
  _, _, kernel_fn = cases.build(spec, nt.stax)
  x = np.random.default_rng(13).standard_normal((8, 32, 32, 3)).astype(np.float32)
  one = kernel_fn(x, None, ('nngp', 'ntk'))
  two = nt.batch(kernel_fn, batch_size=2, device_count=2)(x, None, ('nngp', 'ntk'))
  np.testing.assert_array_equal(two.nngp, one.nngp)
  np.testing.assert_array_equal(two.ntk, one.ntk)
  np.testing.assert_array_equal(nt.batch(kernel_fn, batch_size=2, device_count=2)(x, None, 'ntk'), one.ntk)


@pytest.mark.parametrize('shape', [(28, 28, 1), (32, 32, 1), (24, 20, 3), (30, 30, 3), (20, 32, 1)])
def test_embedded_sizes_fused_path_vs_oracle(nt, shape):
  """Round 2 widening: any H x W <= 32 x 32, C in {1, 3}, embedded in the next shear size (EMB kernels).  Myrtle-10
  body (3 + 3 + 3 fused layers, two pools) with a GlobalAvgPool tail, against the oracle and the per-op path."""
  from oracle import ntk_oracle as O
  H, W, C = shape
  spec = cases.myrtle(10, 'gap')
  _, _, kernel_fn = cases.build(spec, nt.stax)
  low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
  assert low.program.path(H, W, C) == 'fused'
  x1 = np.random.default_rng(31).standard_normal((3, H, W, C)).astype(np.float32)
  x2 = np.random.default_rng(32).standard_normal((2, H, W, C)).astype(np.float32)
  ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
  sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
  for x64 in (False, True):
    nt.config.update('enable_x64', x64)
    out = kernel_fn(x1, x2, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
    np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
    np.testing.assert_allclose(kernel_fn(x1, x2, 'nngp'), ref[0], rtol=RTOL[x64])
    sym = kernel_fn(x1, None, ('nngp', 'ntk'))
    _check_sym(sym.nngp, sref[0], x64)
    _check_sym(sym.ntk, sref[1], x64)
  nt.config.update('enable_x64', False)


def test_valid_convs_run_on_the_diagonal_and_stage_kernels(nt):
  """3x3 / 1 / VALID convs: pool-free Flatten nets on the diagonal-column kernels, pooled / GlobalAvgPool
  nets on the stage kernels (the SAME stencil everywhere, epilogues restricted to the shrinking box), against the oracle
  and the per-op path; a pool behind an odd number of VALID convs is not box-aligned and stays on the per-op path."""
  from oracle import ntk_oracle as O
  V = lambda **kw: cases.conv(pad='VALID', **kw)
  nets = [
      # (spec, shapes, expected path)
      (('serial', [V(W=1.3, b=0.1), cases.RELU, V(), ('abrelu', 0.1, 1., False), V(W=1.1, b=0.05), cases.RELU,
                   ('flatten',), ('dense', 1., 0.1)]), [(12, 12, 3), (28, 28, 1), (9, 9, 2)], 'diag'),
      (('serial', [V(W=1.3, b=0.1), cases.RELU, V(), cases.RELU, cases.pool(), V(W=1.1, b=0.2), cases.RELU, V(), cases.RELU,
                   ('gap',), ('dense', 1.2, 0.1)]), [(32, 32, 3), (28, 28, 1), (30, 26, 3)], 'fused'),
      # four layers: two chunks with a STORE boundary in between (box origin 3 entering the second chunk)
      (('serial', [V(W=1.2, b=0.1), cases.RELU] + [V(), cases.RELU] * 3 + [('gap',), ('dense', 1., 0.)]),
       [(16, 16, 3), (20, 14, 1)], 'fused'),
      # tail pools + Flatten at 1x1: 8 -> 6 -> 4 -pool-> 2 -pool-> 1
      (('serial', [V(b=0.1), cases.RELU, V(), cases.RELU, cases.pool(), cases.pool(), ('flatten',), ('dense', 1., 0.1)]),
       [(8, 8, 3)], 'fused'),
      # pool behind three VALID convs: origin 3 is odd
      (('serial', [V(), cases.RELU] * 3 + [cases.pool(), V(), cases.RELU, ('gap',), ('dense', 1., 0.)]), [(16, 16, 3)],
       'generic'),
  ]
  for spec, shapes, path in nets:
    _, _, kernel_fn = cases.build(spec, nt.stax)
    low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
    for shape in shapes:
      assert low.program.path(*shape) == path, (shape, low.program.path(*shape), path)
      x1 = np.random.default_rng(51).standard_normal((3,) + shape).astype(np.float32)
      x2 = np.random.default_rng(52).standard_normal((2,) + shape).astype(np.float32)
      ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
      sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
      for x64 in (False, True):
        nt.config.update('enable_x64', x64)
        out = kernel_fn(x1, x2, ('nngp', 'ntk'))
        np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64], err_msg=f'{path} {shape}')
        np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64], err_msg=f'{path} {shape}')
        np.testing.assert_allclose(kernel_fn(x1, x2, 'nngp'), ref[0], rtol=RTOL[x64])
        sym = kernel_fn(x1, None, ('nngp', 'ntk'))
        if path == 'generic':   # full square, pair (j, i) sums the transposed tensor in another order: no exact symmetry
          off = ~np.eye(3, dtype=bool)
          np.testing.assert_allclose(sym.ntk[off], sref[1][off], rtol=RTOL[x64])
          np.testing.assert_allclose(np.diag(sym.ntk), np.diag(sref[1]), rtol=RTOL_DUP[x64])
          continue
        _check_sym(sym.nngp, sref[0], x64)
        _check_sym(sym.ntk, sref[1], x64)
  nt.config.update('enable_x64', False)


@pytest.mark.parametrize('shape', [(32, 32, 4), (16, 16, 2), (28, 28, 2), (20, 24, 5)])
def test_any_channel_count_on_the_fused_path(nt, shape):
  """Channel counts without a FROM_X instantiation (C not in {1, 3}): `k_input_shear` writes the sheared input covariance and
  the first stage LOADs it (ntk = 0) -- packed / scalar kernels on 32 / 16 px, the EMB family on other sizes -- against the
  oracle; duplicate pairs keep their exact diagonal (same roundings as k_qmaps)."""
  from oracle import ntk_oracle as O
  H, W, C = shape
  spec = cases.myrtle(5, 'gap') if min(H, W) >= 20 else ('serial', [cases.conv(W=1.3, b=0.1), cases.RELU, cases.conv(), cases.RELU,
                                                                    cases.pool(), cases.conv(), cases.RELU, ('gap',),
                                                                    ('dense', 1.1, 0.1)])
  _, _, kernel_fn = cases.build(spec, nt.stax)
  low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
  assert low.program.path(H, W, C) == 'fused'
  x1 = np.random.default_rng(61).standard_normal((3, H, W, C)).astype(np.float32)
  x2 = np.random.default_rng(62).standard_normal((2, H, W, C)).astype(np.float32)
  ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
  sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
  for x64 in (False, True):
    nt.config.update('enable_x64', x64)
    out = kernel_fn(x1, x2, ('nngp', 'ntk'))
    np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
    np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
    np.testing.assert_allclose(kernel_fn(x1, x2, 'nngp'), ref[0], rtol=RTOL[x64])
    sym = kernel_fn(x1, None, ('nngp', 'ntk'))
    _check_sym(sym.nngp, sref[0], x64)
    _check_sym(sym.ntk, sref[1], x64)
    dup = kernel_fn(x1, x1.copy(), 'ntk')
    np.testing.assert_array_equal(np.diag(dup), np.diag(sym.ntk))
  nt.config.update('enable_x64', False)


def test_grey_erf_network_runs_on_the_fused_path_through_the_prepass(nt):
  """Grey shear-size inputs of networks the ABRelu-only EMB family cannot run (Erf): k_input_shear + the Erf families."""
  from oracle import ntk_oracle as O
  erf = ('erf', 1., 1.1, 0.1)
  spec = ('serial', [cases.conv(W=1.2, b=0.1), erf, cases.conv(), erf, cases.pool(), cases.conv(), cases.RELU, ('gap',),
                     ('dense', 1., 0.1)])
  _, _, kernel_fn = cases.build(spec, nt.stax)
  low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
  for shape in ((32, 32, 1), (16, 16, 2)):
    assert low.program.path(*shape) == 'fused'
    x1 = np.random.default_rng(71).standard_normal((3,) + shape).astype(np.float32)
    x2 = np.random.default_rng(72).standard_normal((2,) + shape).astype(np.float32)
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    for x64 in (False, True):
      nt.config.update('enable_x64', x64)
      out = kernel_fn(x1, x2, ('nngp', 'ntk'))
      np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
      np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
  nt.config.update('enable_x64', False)


def test_embedded_sizes_with_erf_gelu_sin_run_on_the_fused_path(nt):
  """The EMB family with the general activation code (fused_*_emb_gen.cu): MNIST-sized / non-square inputs and VALID stacks
  whose stages contain Erf / Gelu / Sin layers, against the oracle."""
  from oracle import ntk_oracle as O
  erf = ('erf', 1., 1.1, 0.1)
  V = lambda **kw: cases.conv(pad='VALID', **kw)
  nets = [
      (('serial', [cases.conv(W=1.2, b=0.1), erf, cases.conv(), erf, cases.pool(), cases.conv(), cases.RELU, ('gap',),
                   ('dense', 1., 0.1)]), [(28, 28, 1), (20, 24, 3)]),
      (('serial', [cases.conv(W=1.2, b=0.1), ('gelu',), cases.conv(W=1.1, b=0.), ('sin', 1.1, 0.7, 0.2), ('gap',),
                   ('dense', 1., 0.1)]), [(12, 12, 3), (28, 28, 1)]),
      (('serial', [V(W=1.2, b=0.1), erf, V(), ('gelu',), ('gap',), ('dense', 1., 0.)]), [(16, 16, 3), (14, 10, 2)]),
  ]
  for spec, shapes in nets:
    _, _, kernel_fn = cases.build(spec, nt.stax)
    low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
    for shape in shapes:
      assert low.program.path(*shape) == 'fused', shape
      x1 = np.random.default_rng(81).standard_normal((3,) + shape).astype(np.float32)
      x2 = np.random.default_rng(82).standard_normal((2,) + shape).astype(np.float32)
      ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
      for x64 in (False, True):
        nt.config.update('enable_x64', x64)
        out = kernel_fn(x1, x2, ('nngp', 'ntk'))
        np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64], err_msg=f'{shape}')
        np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64], err_msg=f'{shape}')
  nt.config.update('enable_x64', False)


def test_edge_networks_on_the_fused_path(nt):
  """Corners of the round-2 widening: a single-stage network behind the channel pre-pass (the boundary buffers exist only
  for the pre-pass), a single VALID layer, pre-pass + VALID + embedded size, an Erf stage behind the pre-pass, four layers
  (two chunks) at C = 5 -- cross, symmetric and nngp-only calls against the oracle."""
  from oracle import ntk_oracle as O
  V = lambda **kw: cases.conv(pad='VALID', **kw)
  nets = [
      (('serial', [cases.conv(W=1.2, b=0.1), cases.RELU, ('gap',), ('dense', 1., 0.1)]), (16, 16, 4)),
      (('serial', [V(W=1.2, b=0.1), cases.RELU, ('gap',)]), (8, 8, 3)),
      (('serial', [V(), cases.RELU, V(W=1.1, b=0.2), cases.RELU, ('gap',), ('dense', 1., 0.)]), (12, 12, 2)),
      (('serial', [cases.conv(), ('erf', 1., 1., 0.), ('gap',)]), (32, 32, 2)),
      (('serial', [cases.conv(), cases.RELU] * 4 + [('gap',)]), (16, 16, 5)),
  ]
  for spec, shape in nets:
    _, _, kernel_fn = cases.build(spec, nt.stax)
    low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
    assert low.program.path(*shape) == 'fused', shape
    x1 = np.random.default_rng(5).standard_normal((3,) + shape).astype(np.float32)
    x2 = np.random.default_rng(6).standard_normal((2,) + shape).astype(np.float32)
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
    for x64 in (False, True):
      nt.config.update('enable_x64', x64)
      out = kernel_fn(x1, x2, ('nngp', 'ntk'))
      np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64], err_msg=f'{shape}')
      np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64], err_msg=f'{shape}')
      np.testing.assert_allclose(kernel_fn(x1, x2, 'nngp'), ref[0], rtol=RTOL[x64])
      sym = kernel_fn(x1, None, ('nngp', 'ntk'))
      _check_sym(sym.nngp, sref[0], x64)
      _check_sym(sym.ntk, sref[1], x64)
  nt.config.update('enable_x64', False)


def test_sum_pools_on_the_fused_kernels(nt):
  """SumPool / GlobalSumPool (linear.py:1503, 1674) are epilogue scales of the fused kernels."""
  from oracle import ntk_oracle as O
  spec = ('serial', [cases.conv(W=1.2, b=0.1), cases.RELU, cases.conv(), cases.RELU,
                     ('sumpool', (2, 2), (2, 2), 'VALID'), cases.conv(), cases.RELU, ('sumpool', (2, 2), (2, 2), 'SAME'),
                     ('gsp',), ('dense', 1.1, 0.2)])
  _, _, kernel_fn = cases.build(spec, nt.stax)
  for shape in ((32, 32, 3), (28, 28, 1)):
    low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
    assert low.program.path(*shape) == 'fused'
    x1 = np.random.default_rng(41).standard_normal((2,) + shape).astype(np.float32)
    x2 = np.random.default_rng(42).standard_normal((3,) + shape).astype(np.float32)
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    for x64 in (False, True):
      nt.config.update('enable_x64', x64)
      out = kernel_fn(x1, x2, ('nngp', 'ntk'))
      np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64])
      np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64])
  nt.config.update('enable_x64', False)


@pytest.mark.parametrize('size', [32, 16])
def test_gelu_sin_rbf_in_the_fused_stage_kernels(nt, size):
  """Gelu / Sin / Cos / Rbf closed forms (elementwise.py:195-400) inside the fused stage kernels
  (general family), alone and mixed with Relu / Erf in one stage, against the oracle and the per-op path."""
  from oracle import ntk_oracle as O
  specs = {
      'gelu': ('serial', [cases.conv(W=1.2, b=0.1), ('gelu',), cases.conv(W=1.1, b=0.), ('gelu',), cases.pool(),
                          cases.conv(), ('gelu',), ('gap',), ('dense', 1., 0.1)]),
      'sin_cos_rbf': ('serial', [cases.conv(W=1.1, b=0.2), ('sin', 1.2, 0.7, 0.3), cases.conv(W=1., b=0.1),
                                 ('cos', 0.9, 1.1, 0.2), cases.pool(), cases.conv(W=1., b=0.1), ('rbf', 0.5),
                                 ('gap',), ('dense', 1., 0.)]),
      'mixed': ('serial', [cases.conv(W=1.3, b=0.1), cases.RELU, cases.conv(), ('gelu',), cases.conv(), ('erf', 1., 1., 0.),
                           cases.pool(), cases.conv(), ('rbf', 0.8), cases.conv(), cases.RELU, ('gap',), ('dense', 1.1, 0.)]),
  }
  for name, spec in specs.items():
    _, _, kernel_fn = cases.build(spec, nt.stax)
    low = nt.stax._lowered(nt.stax._strip(kernel_fn._spec), False, False, True)
    assert low.program.path(size, size, 3) == 'fused', name
    x1 = np.random.default_rng(51).standard_normal((3, size, size, 3)).astype(np.float32)
    x2 = np.random.default_rng(52).standard_normal((2, size, size, 3)).astype(np.float32)
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'))
    sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'))
    for x64 in (False, True):
      nt.config.update('enable_x64', x64)
      out = kernel_fn(x1, x2, ('nngp', 'ntk'))
      np.testing.assert_allclose(out.nngp, ref[0], rtol=RTOL[x64], err_msg=name)
      np.testing.assert_allclose(out.ntk, ref[1], rtol=RTOL[x64], err_msg=name)
      sym = kernel_fn(x1, None, ('nngp', 'ntk'))
      _check_sym(sym.nngp, sref[0], x64)
      _check_sym(sym.ntk, sref[1], x64)
  nt.config.update('enable_x64', False)
