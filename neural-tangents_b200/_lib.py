"""ctypes binding of libntk_b200.so (the C-ABI declared in include/ntk_b200.h).

Fails loudly: if the shared library is missing or no CUDA device is usable,
every compute entry point raises — there is no CPU or PyTorch fallback.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NTK_B200_LIB') or os.path.join(_HERE, 'libntk_b200.so')  # override: A/B builds (profiles/)

NTK_F32, NTK_F64 = 0, 1
(OP_DENSE, OP_CONV, OP_ABRELU, OP_ERF, OP_AVGPOOL, OP_GAP, OP_FLATTEN, OP_FANINSUM, OP_IDENTITY, OP_GELU, OP_SIN,
 OP_RBF, OP_LAYERNORM) = range(1, 14)
PAD = {'VALID': 0, 'SAME': 1, 'CIRCULAR': 2}
NTK_NONE, NTK_ZERO, NTK_TENSOR = 0, 1, 2
FLAG_NTK, FLAG_NO_FUSION, FLAG_WANT_COV, FLAG_PER_LAYER, FLAG_FULL_SQUARE, FLAG_UPPER_ONLY = 1, 2, 4, 8, 16, 32
COMM_ID_BYTES = 128

E_INVAL, E_CUDA, E_NOMEM, E_NOTGAUSSIAN, E_UNSUPPORTED, E_SHAPE = -1, -2, -3, -4, -5, -6
PATH_NAMES = ('generic', 'fused', 'res', 'diag', 'fcn')   # NTK_PATH_* of include/ntk_b200.h

# every symbol include/ntk_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = (
    'ntk_abi_version', 'ntk_last_error', 'ntk_device_count', 'ntk_program_create',
    'ntk_program_destroy', 'ntk_program_output_shape', 'ntk_program_path', 'ntk_context_create',
    'ntk_context_destroy',
    'ntk_context_synchronize', 'ntk_context_stream', 'ntk_context_launch_count',
    'ntk_context_set_profiling', 'ntk_context_profile', 'ntk_gram_host',
    'ntk_gram_device', 'ntk_apply_host', 'ntk_workspace_bytes', 'ntk_device_malloc',
    'ntk_device_free', 'ntk_memcpy_h2d', 'ntk_memcpy_d2h',
    # ABI version 2
    'ntk_gram_device_on_stream', 'ntk_apply_device', 'ntk_comm_unique_id', 'ntk_comm_create',
    'ntk_comm_destroy', 'ntk_comm_rank', 'ntk_comm_world', 'ntk_comm_nccl_version', 'ntk_comm_broadcast',
    'ntk_comm_all_gather', 'ntk_sym_assemble', 'ntk_memset_async', 'ntk_context_device', 'ntk_host_alloc',
    'ntk_host_free', 'ntk_event_create', 'ntk_event_record', 'ntk_event_elapsed_ms', 'ntk_event_destroy',
    'ntk_stream_create', 'ntk_stream_synchronize', 'ntk_stream_query', 'ntk_stream_destroy',
    'ntk_chol_factor', 'ntk_chol_info', 'ntk_chol_solve', 'ntk_matmul_f64', 'ntk_chol_factor_ptr',
    'ntk_chol_destroy',
    'ntk_eigh_compute', 'ntk_eigh_info', 'ntk_eigh_values', 'ntk_eigh_values_ptr', 'ntk_eigh_vectors_ptr',
    'ntk_eigh_vectors_t_ptr', 'ntk_eigh_destroy')


class NtkOp(ctypes.Structure):
  _fields_ = [('kind', ctypes.c_int32), ('src', ctypes.c_int32), ('src2', ctypes.c_int32),
              ('dst', ctypes.c_int32), ('i', ctypes.c_int32 * 6), ('f', ctypes.c_double * 4)]


class NtkState(ctypes.Structure):
  _fields_ = [('nngp', ctypes.c_void_p), ('ntk', ctypes.c_void_p), ('cov1', ctypes.c_void_p),
              ('cov2', ctypes.c_void_p), ('n1', ctypes.c_int32), ('n2', ctypes.c_int32),
              ('H', ctypes.c_int32), ('W', ctypes.c_int32), ('ntk_mode', ctypes.c_int32),
              ('is_gaussian', ctypes.c_int32)]


class NtkError(RuntimeError):
  def __init__(self, code, msg):
    super().__init__(f'ntk_b200 error {code}: {msg}')
    self.code = code
    self.msg = msg


_lib = None
_lock = threading.Lock()


def _prefer_bundled_nccl():
  """libntk_b200.so resolves NCCL with dlopen on the first `ntk_comm_*` call (NTK_B200_NCCL_LIB, then
  libnccl.so.2 on the loader path).  A process can hold only one libnccl.so.2 (the loader de-duplicates by
  SONAME), so when the Python environment ships a newer NCCL wheel (`nvidia-nccl-cu12`, the one other
  frameworks in the same process were built against) that copy is named explicitly."""
  if os.environ.get('NTK_B200_NCCL_LIB'):
    return
  import importlib.util
  try:
    spec = importlib.util.find_spec('nvidia.nccl')
  except (ImportError, ValueError):
    spec = None
  for base in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
    cand = os.path.join(base, 'lib', 'libnccl.so.2')
    if os.path.exists(cand):
      os.environ['NTK_B200_NCCL_LIB'] = cand
      return


def load():
  """Loads the shared library (once).  Raises if it has not been built."""
  global _lib
  with _lock:
    if _lib is not None:
      return _lib
    if not os.path.exists(LIB_PATH):
      raise ImportError(
          f'{LIB_PATH} not found: build the CUDA library first '
          '(`python -c "import __graft_entry__ as g; g.build()"` or `make -C neural-tangents_b200/csrc`). '
          'neural_tangents_b200 has no CPU fallback.')
    _prefer_bundled_nccl()
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_size_t
    P = ctypes.POINTER
    lib.ntk_abi_version.restype = ctypes.c_int
    lib.ntk_last_error.restype = ctypes.c_char_p
    lib.ntk_device_count.argtypes = [P(ctypes.c_int)]
    lib.ntk_program_create.argtypes = [P(NtkOp), i32, i32, i32, P(vp)]
    lib.ntk_program_destroy.argtypes = [vp]
    lib.ntk_program_destroy.restype = None
    lib.ntk_program_output_shape.argtypes = [vp, i32, i32, i32, P(i32), P(i32), P(i32)]
    lib.ntk_program_path.argtypes = [vp, i32, i32, i32, i32, u32, P(i32)]
    lib.ntk_context_create.argtypes = [i32, sz, P(vp)]
    lib.ntk_context_destroy.argtypes = [vp]
    lib.ntk_context_destroy.restype = None
    lib.ntk_context_synchronize.argtypes = [vp]
    lib.ntk_context_stream.argtypes = [vp]
    lib.ntk_context_stream.restype = vp
    lib.ntk_context_launch_count.argtypes = [vp]
    lib.ntk_context_launch_count.restype = i64
    lib.ntk_context_set_profiling.argtypes = [vp, i32]
    lib.ntk_context_profile.argtypes = [vp, i32, P(ctypes.c_double), P(i64), P(i64)]
    gram_args = [vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, u32, vp, vp, i64, vp, vp]
    lib.ntk_gram_host.argtypes = gram_args
    lib.ntk_gram_device.argtypes = gram_args
    lib.ntk_apply_host.argtypes = [vp, vp, i32, P(NtkState), P(NtkState)]
    lib.ntk_workspace_bytes.argtypes = [vp, i32, i32, i32, i32, i32, i32, u32, P(sz)]
    lib.ntk_device_malloc.argtypes = [i32, sz, P(vp)]
    lib.ntk_device_free.argtypes = [i32, vp]
    lib.ntk_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    lib.ntk_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    lib.ntk_gram_device_on_stream.argtypes = gram_args + [vp]
    lib.ntk_apply_device.argtypes = [vp, vp, i32, P(NtkState), P(NtkState), vp]
    lib.ntk_comm_unique_id.argtypes = [vp]
    lib.ntk_comm_create.argtypes = [vp, vp, i32, i32, P(vp)]
    lib.ntk_comm_destroy.argtypes = [vp]
    lib.ntk_comm_destroy.restype = None
    lib.ntk_comm_rank.argtypes = [vp]
    lib.ntk_comm_world.argtypes = [vp]
    lib.ntk_comm_nccl_version.argtypes = [P(ctypes.c_int)]
    lib.ntk_comm_broadcast.argtypes = [vp, vp, sz, i32]
    lib.ntk_comm_all_gather.argtypes = [vp, vp, vp, sz]
    lib.ntk_sym_assemble.argtypes = [vp, i32, vp, i64, vp, i32, vp, i64]
    lib.ntk_memset_async.argtypes = [vp, vp, i32, sz]
    lib.ntk_context_device.argtypes = [vp]
    lib.ntk_host_alloc.argtypes = [sz, P(vp)]
    lib.ntk_host_free.argtypes = [vp]
    lib.ntk_event_create.argtypes = [P(vp)]
    lib.ntk_event_record.argtypes = [vp, vp]
    lib.ntk_event_elapsed_ms.argtypes = [vp, vp, P(ctypes.c_float)]
    lib.ntk_event_destroy.argtypes = [vp]
    lib.ntk_event_destroy.restype = None
    lib.ntk_stream_create.argtypes = [i32, P(vp)]
    lib.ntk_stream_synchronize.argtypes = [vp]
    lib.ntk_stream_query.argtypes = [vp, P(i32)]
    lib.ntk_stream_destroy.argtypes = [vp]
    lib.ntk_stream_destroy.restype = None
    lib.ntk_chol_factor.argtypes = [vp, i32, vp, i32, i64, ctypes.c_double, i32, P(vp)]
    lib.ntk_chol_info.argtypes = [vp, vp, P(i32)]
    lib.ntk_chol_solve.argtypes = [vp, vp, vp, i32, i64, vp]
    lib.ntk_matmul_f64.argtypes = [vp, i32, vp, i32, i32, i64, vp, i32, i64, vp, i64]
    lib.ntk_chol_factor_ptr.argtypes = [vp]
    lib.ntk_chol_factor_ptr.restype = vp
    lib.ntk_chol_destroy.argtypes = [vp]
    lib.ntk_chol_destroy.restype = None
    lib.ntk_eigh_compute.argtypes = [vp, i32, vp, i32, i64, ctypes.c_double, i32, i32, ctypes.c_double, P(vp)]
    lib.ntk_eigh_info.argtypes = [vp, P(i32), P(ctypes.c_double)]
    lib.ntk_eigh_values.argtypes = [vp, vp]
    for f in (lib.ntk_eigh_values_ptr, lib.ntk_eigh_vectors_ptr, lib.ntk_eigh_vectors_t_ptr):
      f.argtypes = [vp]
      f.restype = vp
    lib.ntk_eigh_destroy.argtypes = [vp]
    lib.ntk_eigh_destroy.restype = None
    _lib = lib
    return lib


def check(status):
  if status != 0:
    msg = load().ntk_last_error().decode('utf-8', 'replace')
    if status in (E_NOTGAUSSIAN, E_SHAPE, E_INVAL):
      raise ValueError(msg)
    if status == E_UNSUPPORTED:
      raise NotImplementedError(msg)
    if status == E_NOMEM:
      raise MemoryError(msg)
    raise NtkError(status, msg)


def dtype_code(dtype):
  dtype = np.dtype(dtype)
  if dtype == np.float32:
    return NTK_F32
  if dtype == np.float64:
    return NTK_F64
  raise TypeError(f'unsupported dtype {dtype}')


class Program:
  """Owns an `ntk_program_t`."""

  def __init__(self, ops, n_slots, out_slot):
    lib = load()
    arr = (NtkOp * max(len(ops), 1))()
    for k, (kind, src, src2, dst, ints, floats) in enumerate(ops):
      o = arr[k]
      o.kind, o.src, o.src2, o.dst = kind, src, src2, dst
      for t in range(6):
        o.i[t] = int(ints[t]) if t < len(ints) else 0
      for t in range(4):
        o.f[t] = float(floats[t]) if t < len(floats) else 0.0
    self._h = ctypes.c_void_p()
    check(lib.ntk_program_create(arr, len(ops), n_slots, out_slot, ctypes.byref(self._h)))
    self._lib = lib

  @property
  def handle(self):
    return self._h

  def output_shape(self, H, W, in_is_gaussian=False):
    oh, ow, og = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    check(self._lib.ntk_program_output_shape(self._h, H, W, int(in_is_gaussian), ctypes.byref(oh), ctypes.byref(ow),
                                              ctypes.byref(og)))
    return oh.value, ow.value, bool(og.value)

  def path(self, H, W, C, x64=False, flags=0):
    """Kernel family `ntk_gram_*` runs this program on: 'generic', 'fused', 'res', 'diag' or 'fcn'."""
    out = ctypes.c_int32()
    check(self._lib.ntk_program_path(self._h, NTK_F64 if x64 else NTK_F32, H, W, C, flags, ctypes.byref(out)))
    return PATH_NAMES[out.value]

  def __del__(self):
    try:
      if self._h:
        self._lib.ntk_program_destroy(self._h)
    except Exception:
      pass


class Context:
  """Owns an `ntk_context_t` (one per host thread x GPU)."""

  def __init__(self, device=0, workspace_bytes=0):
    lib = load()
    self._h = ctypes.c_void_p()
    check(lib.ntk_context_create(device, workspace_bytes, ctypes.byref(self._h)))
    self._lib = lib
    self.device = device
    # one context per GPU is shared by every host thread that computes on it (workspace, stream and staging
    # buffers are per context): calls are serialised with this lock
    self.lock = threading.RLock()
    self._pool, self._sizes, self._pooled = {}, {}, 0

  @property
  def handle(self):
    return self._h

  # ---- device memory / events through the C-ABI (a host language needs no CUDA binding) ----
  # cudaMalloc / cudaFree cost 0.1 - 1 ms each and cudaFree synchronises the device: blocks are recycled through a
  # per-context pool (same-size reuse; at most _POOL_CAP bytes are kept).  Work on a recycled block is ordered by
  # the context stream, on which every user of these blocks runs.
  _POOL_CAP = 8 << 30

  def malloc(self, nbytes):
    size = (max(int(nbytes), 1) + 255) & ~255
    with self.lock:
      free = self._pool.get(size)
      if free:
        self._pooled -= size
        ptr = free.pop()
        self._sizes[ptr] = size
        return ptr
    p = ctypes.c_void_p()
    check(self._lib.ntk_device_malloc(self.device, size, ctypes.byref(p)))
    with self.lock:
      self._sizes[p.value] = size
    return p.value

  def free(self, ptr):
    if not ptr:
      return
    with self.lock:
      size = self._sizes.pop(ptr, None)
      if size is not None and self._pooled + size <= self._POOL_CAP and self._h:
        self._pool.setdefault(size, []).append(ptr)
        self._pooled += size
        return
    check(self._lib.ntk_device_free(self.device, ctypes.c_void_p(ptr)))

  def trim(self):
    """Returns the pooled device blocks to the driver."""
    with self.lock:
      blocks = [p for lst in self._pool.values() for p in lst]
      self._pool.clear()
      self._pooled = 0
    for p in blocks:
      self._lib.ntk_device_free(self.device, ctypes.c_void_p(p))

  def h2d(self, dst_ptr, arr):
    arr = np.ascontiguousarray(arr)
    check(self._lib.ntk_memcpy_h2d(self._h, ctypes.c_void_p(dst_ptr), _ptr(arr), arr.nbytes))
    return arr   # keep alive until the stream has consumed it (callers synchronise)

  def d2h(self, arr, src_ptr):
    check(self._lib.ntk_memcpy_d2h(self._h, _ptr(arr), ctypes.c_void_p(src_ptr), arr.nbytes))
    return arr

  def memset(self, ptr, value, nbytes):
    check(self._lib.ntk_memset_async(self._h, ctypes.c_void_p(ptr), value, nbytes))

  def synchronize(self):
    check(self._lib.ntk_context_synchronize(self._h))

  @property
  def stream(self):
    return self._lib.ntk_context_stream(self._h)

  @property
  def launch_count(self):
    return int(self._lib.ntk_context_launch_count(self._h))

  def set_profiling(self, enabled):
    check(self._lib.ntk_context_set_profiling(self._h, int(bool(enabled))))

  def profile(self, stage):
    """(total_ms, launches, pairs) of fused stage `stage` since profiling was enabled."""
    ms, n, pr = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    check(self._lib.ntk_context_profile(self._h, stage, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(pr)))
    return ms.value, n.value, pr.value

  def close(self):
    if self._h:
      self.synchronize()
      self.trim()
      self._lib.ntk_context_destroy(self._h)
      self._h = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


_contexts = {}
_ctx_lock = threading.Lock()
_tls = threading.local()


class device_scope:
  """`with device_scope(d):` routes this thread's kernel_fn calls to GPU `d`."""

  def __init__(self, device):
    self.device = device

  def __enter__(self):
    self.prev = getattr(_tls, 'device', None)
    _tls.device = self.device

  def __exit__(self, *exc):
    _tls.device = self.prev


def device_count():
  n = ctypes.c_int(0)
  check(load().ntk_device_count(ctypes.byref(n)))
  return n.value


def get_context(device=None):
  """The context of GPU `device` (one per device, created on first use; callers hold `ctx.lock` while they
  run on it).  `nt.batch(device_count=D)` spawns fresh host threads on every call: keying the cache by thread
  would leak a workspace per call."""
  from ._config import config
  if device is None:
    device = getattr(_tls, 'device', None)
  if device is None:
    device = config.device
  with _ctx_lock:
    ctx = _contexts.get(device)
    if ctx is None:
      ctx = Context(device, config.workspace_bytes)
      _contexts[device] = ctx
    return ctx


def close_contexts():
  """Destroys every cached context (frees the device workspaces)."""
  with _ctx_lock:
    for ctx in _contexts.values():
      ctx.close()
    _contexts.clear()


# ---- page-locked host arrays -----------------------------------------------------------------------
# Results of `gram_host` are allocated in pinned memory, so the device -> host copy is a direct DMA at PCIe
# speed (pageable destinations go through the context's staging ring instead).  Blocks are recycled through a
# small pool: cudaMallocHost / cudaFreeHost cost milliseconds and synchronise the device.
_PIN_MIN, _PIN_MAX, _PIN_POOL_CAP = 1 << 16, 1 << 28, 1 << 30
_pin_pool = {}          # rounded size -> [ptr, ...]
_pin_pooled_bytes = 0
_pin_lock = threading.Lock()


def _pin_release(ptr, size):
  global _pin_pooled_bytes
  with _pin_lock:
    if _pin_pooled_bytes + size <= _PIN_POOL_CAP:
      _pin_pool.setdefault(size, []).append(ptr)
      _pin_pooled_bytes += size
      return
  try:
    load().ntk_host_free(ctypes.c_void_p(ptr))
  except Exception:
    pass


def pinned_empty(shape, dtype):
  """`np.empty(shape, dtype)` in page-locked memory (falls back to pageable memory for tiny / huge arrays)."""
  global _pin_pooled_bytes
  dtype = np.dtype(dtype)
  nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
  if nbytes < _PIN_MIN or nbytes > _PIN_MAX:
    return np.empty(shape, dtype)
  size = 1 << (nbytes - 1).bit_length()
  ptr = None
  with _pin_lock:
    free = _pin_pool.get(size)
    if free:
      ptr = free.pop()
      _pin_pooled_bytes -= size
  if ptr is None:
    p = ctypes.c_void_p()
    if load().ntk_host_alloc(size, ctypes.byref(p)) != 0 or not p.value:
      return np.empty(shape, dtype)
    ptr = p.value
  buf = (ctypes.c_char * size).from_address(ptr)
  import weakref
  weakref.finalize(buf, _pin_release, ptr, size)   # `buf` is the base of every view of the array
  return np.frombuffer(buf, dtype=dtype, count=nbytes // dtype.itemsize).reshape(shape)


def pinned_copy(a):
  out = pinned_empty(a.shape, a.dtype)
  np.copyto(out, a)
  return out


def _ptr(a):
  return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def gram_host(ctx, prog, x1, x2, H, W, C, flags, out_h, out_w, want_ntk, want_cov, device_ptrs=False):
  """kernel_fn on raw inputs; returns dict of numpy arrays in canonical layout."""
  lib = load()
  dt = dtype_code(x1.dtype)
  n1 = x1.shape[0]
  n2 = n1 if x2 is None else x2.shape[0]
  sp = (out_h, out_h, out_w, out_w) if out_h > 0 else ()
  nngp = pinned_empty((n1, n2) + sp, x1.dtype)
  ntk = pinned_empty((n1, n2) + sp, x1.dtype) if want_ntk else None
  cov1 = np.empty((n1,) + sp, x1.dtype) if want_cov else None
  cov2 = np.empty((n2,) + sp, x1.dtype) if (want_cov and x2 is not None) else None
  f = flags | (FLAG_NTK if want_ntk else 0) | (FLAG_WANT_COV if want_cov else 0)
  with ctx.lock:
    check(lib.ntk_gram_host(ctx.handle, prog.handle, dt, _ptr(x1), n1, _ptr(x2), n2, H, W, C, f,
                            _ptr(nngp), _ptr(ntk), n2, _ptr(cov1), _ptr(cov2)))
  return dict(nngp=nngp, ntk=ntk, cov1=cov1, cov2=cov2)


def gram_device(ctx, prog, dtype, x1_ptr, n1, x2_ptr, n2, H, W, C, flags, nngp_ptr, ntk_ptr, ld, stream=None):
  """Device-pointer entry; pointers are ints.  Asynchronous on the context stream, or on `stream` (a
  cudaStream_t as an int: the caller-stream entry `ntk_gram_device_on_stream`)."""
  lib = load()
  f = flags | (FLAG_NTK if ntk_ptr else 0)
  with ctx.lock:
    if stream is None:
      check(lib.ntk_gram_device(ctx.handle, prog.handle, dtype_code(dtype), x1_ptr, n1, x2_ptr, n2, H, W,
                                C, f, nngp_ptr, ntk_ptr, ld, None, None))
    else:
      check(lib.ntk_gram_device_on_stream(ctx.handle, prog.handle, dtype_code(dtype), x1_ptr, n1, x2_ptr, n2, H, W,
                                          C, f, nngp_ptr, ntk_ptr, ld, None, None, ctypes.c_void_p(stream)))


def apply_host(ctx, prog, dtype, nngp, ntk, cov1, cov2, H, W, ntk_mode, is_gaussian, out_h, out_w):
  """Kernel-in / Kernel-out on canonical-layout numpy arrays."""
  lib = load()
  dt = dtype_code(dtype)
  n1, n2 = nngp.shape[0], nngp.shape[1]
  sp = (out_h, out_h, out_w, out_w) if out_h > 0 else ()
  o_nngp = np.empty((n1, n2) + sp, dtype)
  o_ntk = np.empty((n1, n2) + sp, dtype) if ntk_mode != NTK_NONE else None
  o_cov1 = np.empty((n1,) + sp, dtype)
  o_cov2 = np.empty((n2,) + sp, dtype) if cov2 is not None else None
  sin = NtkState(_ptr(nngp), _ptr(ntk) if ntk_mode == NTK_TENSOR else None, _ptr(cov1), _ptr(cov2),
                 n1, n2, H, W, ntk_mode, int(is_gaussian))
  sout = NtkState(_ptr(o_nngp), _ptr(o_ntk), _ptr(o_cov1), _ptr(o_cov2), n1, n2, out_h, out_w, 0, 0)
  with ctx.lock:
    check(lib.ntk_apply_host(ctx.handle, prog.handle, dt, ctypes.byref(sin), ctypes.byref(sout)))
  if sout.ntk_mode == NTK_ZERO and o_ntk is not None:
    o_ntk = np.zeros((), dtype)
  return dict(nngp=o_nngp, ntk=o_ntk, cov1=o_cov1, cov2=o_cov2, ntk_mode=sout.ntk_mode,
              is_gaussian=bool(sout.is_gaussian))


def apply_device(ctx, prog, dtype, n1, n2, H, W, ntk_mode, is_gaussian, in_ptrs, out_ptrs, out_h, out_w,
                 stream=None):
  """Kernel-in / Kernel-out on DEVICE pointers (ints; `in_ptrs` / `out_ptrs` = (nngp, ntk, cov1, cov2), 0 / None
  for absent tensors).  Asynchronous on the context stream or on `stream`.  Returns (ntk_mode, is_gaussian)."""
  lib = load()
  v = lambda p: ctypes.c_void_p(p) if p else None
  sin = NtkState(v(in_ptrs[0]), v(in_ptrs[1]) if ntk_mode == NTK_TENSOR else None, v(in_ptrs[2]), v(in_ptrs[3]),
                 n1, n2, H, W, ntk_mode, int(is_gaussian))
  sout = NtkState(v(out_ptrs[0]), v(out_ptrs[1]), v(out_ptrs[2]), v(out_ptrs[3]), n1, n2, out_h, out_w, 0, 0)
  with ctx.lock:
    check(lib.ntk_apply_device(ctx.handle, prog.handle, dtype_code(dtype), ctypes.byref(sin), ctypes.byref(sout),
                               ctypes.c_void_p(stream) if stream else None))
  return sout.ntk_mode, bool(sout.is_gaussian)


class Event:
  """CUDA event on a context stream (`ntk_event_*`)."""

  def __init__(self):
    self._lib = load()
    self._h = ctypes.c_void_p()
    check(self._lib.ntk_event_create(ctypes.byref(self._h)))

  def record(self, ctx):
    check(self._lib.ntk_event_record(ctx.handle, self._h))

  def elapsed_ms(self, stop):
    ms = ctypes.c_float()
    check(self._lib.ntk_event_elapsed_ms(self._h, stop._h, ctypes.byref(ms)))
    return ms.value

  def __del__(self):
    try:
      self._lib.ntk_event_destroy(self._h)
    except Exception:
      pass


class Comm:
  """NCCL communicator bound to a context (`ntk_comm_*`): one rank per GPU."""

  def __init__(self, ctx, unique_id: bytes, rank: int, world: int):
    self._lib = load()
    self.ctx = ctx
    self._h = ctypes.c_void_p()
    if len(unique_id) != COMM_ID_BYTES:
      raise ValueError(f'unique id must be {COMM_ID_BYTES} bytes')
    buf = ctypes.create_string_buffer(unique_id, COMM_ID_BYTES)
    check(self._lib.ntk_comm_create(ctx.handle, buf, rank, world, ctypes.byref(self._h)))
    self.rank, self.world = rank, world

  @staticmethod
  def unique_id() -> bytes:
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    check(load().ntk_comm_unique_id(buf))
    return buf.raw

  @staticmethod
  def nccl_version() -> int:
    v = ctypes.c_int()
    check(load().ntk_comm_nccl_version(ctypes.byref(v)))
    return v.value

  def broadcast(self, dev_ptr, nbytes, root=0):
    check(self._lib.ntk_comm_broadcast(self._h, ctypes.c_void_p(dev_ptr), nbytes, root))

  def all_gather(self, send_ptr, recv_ptr, nbytes_per_rank):
    check(self._lib.ntk_comm_all_gather(self._h, ctypes.c_void_p(send_ptr), ctypes.c_void_p(recv_ptr),
                                        nbytes_per_rank))

  def close(self):
    if self._h:
      self._lib.ntk_comm_destroy(self._h)
      self._h = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


def sym_assemble(ctx, dtype, slabs_ptr, ld_slabs, row_of_ptr, n, out_ptr, ld_out):
  with ctx.lock:
    check(load().ntk_sym_assemble(ctx.handle, dtype_code(dtype), ctypes.c_void_p(slabs_ptr), ld_slabs,
                                  ctypes.c_void_p(row_of_ptr), n, ctypes.c_void_p(out_ptr), ld_out))


class DeviceCholesky:
  """Regularised Cholesky factor of a device-resident symmetric matrix (`ntk_chol_*`): `(K + reg I) = L L^T` in
  float64 on the GPU that holds K.  `solve(b)` takes / returns host arrays of shape [n, ...]; `matmul(a_ptr, ...)`
  multiplies a device matrix with the last solution without leaving the device."""

  def __init__(self, ctx, dtype, k_ptr, n, ld, diag_reg=0., absolute=False):
    self._lib, self.ctx, self.n = load(), ctx, int(n)
    self._h = ctypes.c_void_p()
    with ctx.lock:
      check(self._lib.ntk_chol_factor(ctx.handle, dtype_code(dtype), ctypes.c_void_p(k_ptr), n, ld, float(diag_reg),
                                      int(bool(absolute)), ctypes.byref(self._h)))
      info = ctypes.c_int32()
      check(self._lib.ntk_chol_info(ctx.handle, self._h, ctypes.byref(info)))
    if info.value != 0:
      self.close()
      raise np.linalg.LinAlgError(f'{info.value}-th leading minor of the array is not positive definite')

  def solve_device(self, b_ptr, nrhs, ldb=None):
    """In place on a device [n, nrhs] float64 matrix."""
    work = self.ctx.malloc(self.n * nrhs * 8)
    try:
      with self.ctx.lock:
        check(self._lib.ntk_chol_solve(self.ctx.handle, self._h, ctypes.c_void_p(b_ptr), nrhs, ldb or nrhs,
                                       ctypes.c_void_p(work)))
        self.ctx.synchronize()
    finally:
      self.ctx.free(work)

  def solve(self, b):
    b = np.asarray(b, np.float64)
    b2 = np.ascontiguousarray(b.reshape(self.n, -1))
    d = self.ctx.malloc(b2.nbytes)
    try:
      self.ctx.h2d(d, b2)
      self.solve_device(d, b2.shape[1])
      out = self.ctx.d2h(np.empty_like(b2), d)
    finally:
      self.ctx.free(d)
    return out.reshape(b.shape)

  def matmul(self, a_dtype, a_ptr, m, lda, x):
    """[m, n] device matrix (a_dtype) times the host matrix x [n, nrhs] -> host [m, nrhs] float64."""
    x2 = np.ascontiguousarray(np.asarray(x, np.float64).reshape(self.n, -1))
    nrhs = x2.shape[1]
    dx, do = self.ctx.malloc(x2.nbytes), self.ctx.malloc(m * nrhs * 8)
    try:
      self.ctx.h2d(dx, x2)
      with self.ctx.lock:
        check(self._lib.ntk_matmul_f64(self.ctx.handle, dtype_code(a_dtype), ctypes.c_void_p(a_ptr), m, self.n, lda,
                                       ctypes.c_void_p(dx), nrhs, nrhs, ctypes.c_void_p(do), nrhs))
      return self.ctx.d2h(np.empty((m, nrhs), np.float64), do)
    finally:
      self.ctx.free(dx)
      self.ctx.free(do)

  def factor(self):
    """The lower factor L as a host array (tests)."""
    return self.ctx.d2h(np.empty((self.n, self.n), np.float64), self._lib.ntk_chol_factor_ptr(self._h))

  def close(self):
    if self._h:
      self._lib.ntk_chol_destroy(self._h)
      self._h = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


class DeviceEigh:
  """Eigendecomposition of a regularised device-resident symmetric matrix (`ntk_eigh_*`, parallel Jacobi in float64 on
  the GPU that holds K): `w` ascending on the host, V / V^T stay in HBM.  `project(x)` = V^T x and `expand(z)` = V z
  take / return host matrices [n, k]; `expand_through(a_ptr, ...)` = A (V z) for a device matrix A (K_test_train)."""

  def __init__(self, ctx, dtype, k_ptr, n, ld, diag_reg=0., absolute=False, max_sweeps=0, tol=0.):
    self._lib, self.ctx, self.n = load(), ctx, int(n)
    self._h = ctypes.c_void_p()
    with ctx.lock:
      check(self._lib.ntk_eigh_compute(ctx.handle, dtype_code(dtype), ctypes.c_void_p(k_ptr), n, ld, float(diag_reg),
                                       int(bool(absolute)), int(max_sweeps), float(tol), ctypes.byref(self._h)))
    sweeps, off = ctypes.c_int32(), ctypes.c_double()
    check(self._lib.ntk_eigh_info(self._h, ctypes.byref(sweeps), ctypes.byref(off)))
    self.sweeps, self.off_over_norm = sweeps.value, off.value
    self.w = np.empty(self.n, np.float64)
    check(self._lib.ntk_eigh_values(self._h, self.w.ctypes.data_as(ctypes.c_void_p)))

  def _mm(self, a_dtype, a_ptr, m, k, lda, x):
    x2 = np.ascontiguousarray(np.asarray(x, np.float64).reshape(k, -1))
    nrhs = x2.shape[1]
    dx, do = self.ctx.malloc(x2.nbytes), self.ctx.malloc(m * nrhs * 8)
    try:
      self.ctx.h2d(dx, x2)
      with self.ctx.lock:
        check(self._lib.ntk_matmul_f64(self.ctx.handle, dtype_code(a_dtype), ctypes.c_void_p(a_ptr), m, k, lda,
                                       ctypes.c_void_p(dx), nrhs, nrhs, ctypes.c_void_p(do), nrhs))
      return self.ctx.d2h(np.empty((m, nrhs), np.float64), do)
    finally:
      self.ctx.free(dx)
      self.ctx.free(do)

  def project(self, x):
    return self._mm(np.float64, self._lib.ntk_eigh_vectors_t_ptr(self._h), self.n, self.n, self.n, x)

  def expand(self, z):
    return self._mm(np.float64, self._lib.ntk_eigh_vectors_ptr(self._h), self.n, self.n, self.n, z)

  def expand_through(self, a_dtype, a_ptr, m, lda, z):
    """A (V z) with A [m, n] on the device, z [n, k] on the host -> host [m, k]; V z never leaves the device."""
    z2 = np.ascontiguousarray(np.asarray(z, np.float64).reshape(self.n, -1))
    k = z2.shape[1]
    dz, du, do = self.ctx.malloc(z2.nbytes), self.ctx.malloc(z2.nbytes), self.ctx.malloc(m * k * 8)
    try:
      self.ctx.h2d(dz, z2)
      with self.ctx.lock:
        check(self._lib.ntk_matmul_f64(self.ctx.handle, NTK_F64, ctypes.c_void_p(self._lib.ntk_eigh_vectors_ptr(self._h)),
                                       self.n, self.n, self.n, ctypes.c_void_p(dz), k, k, ctypes.c_void_p(du), k))
        check(self._lib.ntk_matmul_f64(self.ctx.handle, dtype_code(a_dtype), ctypes.c_void_p(a_ptr), m, self.n, lda,
                                       ctypes.c_void_p(du), k, k, ctypes.c_void_p(do), k))
      return self.ctx.d2h(np.empty((m, k), np.float64), do)
    finally:
      for d in (dz, du, do):
        self.ctx.free(d)

  def vectors(self):
    """V as a host array (tests)."""
    return self.ctx.d2h(np.empty((self.n, self.n), np.float64), self._lib.ntk_eigh_vectors_ptr(self._h))

  def close(self):
    if self._h:
      self._lib.ntk_eigh_destroy(self._h)
      self._h = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass
