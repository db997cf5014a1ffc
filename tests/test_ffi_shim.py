"""The XLA-FFI shim (neural-tangents_b200/csrc/ntk_b200_ffi.cc).  jaxlib's header is not installable here, so the
file is compiled against a stand-in for the API subset it uses (tests/ffi_mock/xla/ffi/api/ffi.h; the Ffi::Bind()
chain is type-checked against the handler signature) and the handler body is executed on the GPU through
tests/ffi_mock/harness.cc: XLA-style buffers, a foreign stream, no host synchronisation inside the handler."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, 'neural-tangents_b200', 'csrc', 'ntk_b200_ffi.cc')


def _build(out_dir):
  if shutil.which('g++') is None:
    pytest.skip('g++ not available')
  import __graft_entry__ as g
  g.build()
  so = os.path.join(str(out_dir), 'libntk_b200_ffi_mock.so')
  libdir = os.path.join(ROOT, 'neural-tangents_b200')
  cmd = ['g++', '-std=c++17', '-O1', '-fPIC', '-shared', '-Wall', '-I' + os.path.join(ROOT, 'tests', 'ffi_mock'),
         '-I' + os.path.join(ROOT, 'include'), '-I/usr/local/cuda/include', '-o', so,
         os.path.join(ROOT, 'tests', 'ffi_mock', 'harness.cc'), '-L' + libdir, '-lntk_b200', '-Wl,-rpath,' + libdir]
  r = subprocess.run(cmd, capture_output=True, text=True)
  assert r.returncode == 0, r.stderr[-3000:]
  return so


def test_shim_compiles_against_the_api_stand_in_and_docs_embed_it(tmp_path):
  so = _build(tmp_path)
  syms = subprocess.run(['nm', '-D', so], capture_output=True, text=True).stdout
  assert 'ntk_ffi_mock_gram' in syms and 'NtkGram_mock_symbol' in syms
  src = open(SHIM).read()
  assert 'ffi::PlatformStream<cudaStream_t>' in src and 'ntk_gram_device_on_stream' in src
  assert 'ntk_context_synchronize' not in src.split('#ifdef NTK_B200_HAVE_XLA_FFI')[1]   # no host sync in the handler
  assert src in open(os.path.join(ROOT, 'INTEGRATION.md')).read(), 'INTEGRATION.md §2 must embed the shim verbatim'
  # without any xla header on the include path the file is an empty translation unit
  r = subprocess.run(['g++', '-std=c++17', '-fsyntax-only', '-I' + os.path.join(ROOT, 'include'), SHIM],
                     capture_output=True, text=True)
  assert r.returncode == 0, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize('x64', [False, True])
def test_handler_body_runs_on_a_foreign_stream(tmp_path, x64):
  so = _build(tmp_path)
  import neural_tangents_b200 as nt
  from neural_tangents_b200 import _lib, stax
  nt.config.update('enable_x64', x64)
  try:
    dt = np.float64 if x64 else np.float32
    lib = _lib.load()
    mock = ctypes.CDLL(so)
    i64p = ctypes.POINTER(ctypes.c_int64)
    mock.ntk_ffi_mock_gram.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_void_p, i64p, ctypes.c_int32, ctypes.c_void_p, i64p, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32]
    spec = cases.myrtle(5)
    _, _, kernel_fn = cases.build(spec, stax)
    low = stax._lowered(stax._strip(kernel_fn._spec), False, False, True)
    x1 = np.random.default_rng(1).standard_normal((5, 32, 32, 3)).astype(np.float32).astype(dt)
    x2 = np.random.default_rng(2).standard_normal((4, 32, 32, 3)).astype(np.float32).astype(dt)
    ref = kernel_fn(x1, x2, ('nngp', 'ntk'))
    ctx = _lib.get_context()
    stream = ctypes.c_void_p()
    _lib.check(lib.ntk_stream_create(ctx.device, ctypes.byref(stream)))
    d1, d2 = ctx.malloc(x1.nbytes), ctx.malloc(x2.nbytes)
    ctx.h2d(d1, x1)
    ctx.h2d(d2, x2)
    ctx.synchronize()
    nb = 5 * 4 * np.dtype(dt).itemsize
    dk, dn = ctx.malloc(nb), ctx.malloc(nb)
    err = ctypes.create_string_buffer(512)
    dims = lambda a: (ctypes.c_int64 * a.ndim)(*a.shape)
    rc = mock.ntk_ffi_mock_gram(stream, low.program.handle.value, ctx.handle.value, _lib.FLAG_NTK, int(x64), d1, dims(x1),
                                4, d2, dims(x2), 4, dk, dn, err, 512)
    assert rc == 0, err.value
    _lib.check(lib.ntk_stream_synchronize(stream))
    k = ctx.d2h(np.empty((5, 4), dt), dk)
    t = ctx.d2h(np.empty((5, 4), dt), dn)
    np.testing.assert_array_equal(k, ref.nngp)
    np.testing.assert_array_equal(t, ref.ntk)
    # error path: a 3-D x1 is rejected by the handler, the message crosses the boundary
    bad = (ctypes.c_int64 * 3)(5, 32, 96)
    rc = mock.ntk_ffi_mock_gram(stream, low.program.handle.value, ctx.handle.value, 1, int(x64), d1, bad, 3, d2, dims(x2),
                                4, dk, dn, err, 512)
    assert rc == 1 and b'x1 must be' in err.value
    for p in (d1, d2, dk, dn):
      ctx.free(p)
    lib.ntk_stream_destroy(stream)
  finally:
    nt.config.update('enable_x64', False)
