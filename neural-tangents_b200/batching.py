"""`nt.batch`: block tiling of the x1 x x2 Gram matrix over one or more GPUs.

API and error behaviour follow `_src/batching.py:76-132` (`batch`), `:314-502`
(`_serial`), `:505-644` (`_parallel`) and `:647-679` (batch-size arithmetic).

B200-first differences (results are identical):
  * A device-parallel call gives every GPU a contiguous slab of x1 rows and the
    whole of x2 (row partition, no reduction collective); the reference's
    interleaved per-super-row assignment (`:594-597`) is an artefact of its
    Python loop.  One host thread drives each GPU through its own
    `ntk_context_t`.
  * Results are returned as host (NumPy) arrays, so `store_on_device` is
    accepted for compatibility and has no effect.
  * For a `kernel_fn` produced by `neural_tangents_b200.stax` whose outputs are
    `[n1, n2]` matrices, a slab is handed to the CUDA library in one call and
    tiled there (`ntk_gram_host`), which keeps x2 and its variance maps resident
    instead of re-sending them for every block (`:731-734,779`).
For multi-process launches (one rank per GPU under torchrun) see
`neural_tangents_b200.distributed`.
"""
import math
import threading
import warnings
from typing import Callable

import numpy as np

from . import _lib
from .kernel import Kernel


def _get_n_batches_and_batch_sizes(n1: int, n2: int, batch_size: int, device_count: int):
  """`_src/batching.py:647-679` (same messages)."""
  max_serial_batch_size = math.gcd(n1, n2) // device_count
  n2_batch_size = min(batch_size, max_serial_batch_size)
  if n2_batch_size != batch_size:
    warnings.warn('Batch size is reduced from requested %d to effective %d to fit the dataset.' %
                  (batch_size, n2_batch_size))
  if n2_batch_size <= 0:
    raise ValueError('Dataset is too small for the requested device_count: '
                     f'gcd(n1, n2) = {math.gcd(n1, n2)} < device_count = {device_count}.')
  n1_batch_size = n2_batch_size * device_count
  n1_batches, ragged = divmod(n1, n1_batch_size)
  if ragged:
    msg = ('Number of rows of kernel must divide batch size. Found n1 = {} '
           'and batch size = {}.').format(n1, n1_batch_size)
    if device_count > 1:
      msg += (' Note that device parallelism was detected and so the batch '
              'size was expanded by a factor of {}.'.format(device_count))
    raise ValueError(msg)
  n2_batches, ragged = divmod(n2, n2_batch_size)
  if ragged:
    raise ValueError(('Number of columns of kernel must divide batch '
                      'size. Found n2 = {} '
                      'and batch size = {}').format(n2, n2_batch_size))
  return n1_batches, n1_batch_size, n2_batches, n2_batch_size


def _get_n_per_device(n1: int, device_count: int):
  """`_src/batching.py:560-573`."""
  n1_per_device, ragged = divmod(n1, device_count)
  if n1_per_device and ragged:
    raise ValueError(('Dataset size ({}) must divide number of '
                      'physical devices ({}).').format(n1, device_count))
  elif not n1_per_device:
    device_count = ragged
    n1_per_device = 1
  return n1_per_device, device_count


# ---- stitching of block results (`_flatten_kernel`, `_src/batching.py:194-267`) ----
def _is_nt(x):
  return hasattr(x, '_asdict') and hasattr(x, '_replace')


def _stitch_arrays(blocks, axis_layout):
  """blocks[r][c] arrays; layout '2' = [n1,n2,...], 'row' = [n1,...], 'col' = [n2,...]."""
  if axis_layout == '2':
    return np.concatenate([np.concatenate(row, axis=1) for row in blocks], axis=0)
  if axis_layout == 'row':
    return np.concatenate([row[0] for row in blocks], axis=0)
  return np.concatenate(list(blocks[0]), axis=0)


def _stitch_field(key, vals, x2_is_none, n1, n2):
  v00 = vals[0][0]
  if key in ('nngp', 'ntk'):
    if v00 is None or np.ndim(v00) == 0:
      return v00
    return _stitch_arrays(vals, '2')
  if key == 'cov1':
    return None if v00 is None else _stitch_arrays(vals, 'row')
  if key == 'cov2':
    if x2_is_none or v00 is None:
      return None
    return _stitch_arrays(vals, 'col')
  if key == 'shape1':
    return (n1,) + tuple(v00[1:])
  if key == 'shape2':
    return (n2,) + tuple(v00[1:])
  if key == 'x1_is_x2':
    return bool(all(all(bool(v) for v in row) for row in vals)) if not x2_is_none else True
  if isinstance(v00, np.ndarray) and v00.ndim >= 2:
    return _stitch_arrays(vals, '2')
  return v00


def _stitch(blocks, x2_is_none, n1, n2):
  b00 = blocks[0][0]
  if isinstance(b00, Kernel):
    d = {k: _stitch_field(k, [[b.asdict()[k] for b in row] for row in blocks], x2_is_none, n1, n2)
         for k in b00.asdict()}
    return Kernel(**d)
  if _is_nt(b00):
    d = {k: _stitch_field(k, [[getattr(b, k) for b in row] for row in blocks], x2_is_none, n1, n2)
         for k in b00._fields}
    return type(b00)(**d)
  if isinstance(b00, np.ndarray):
    if b00.ndim == 0:
      return b00
    return _stitch_arrays(blocks, '2')
  raise TypeError(f'Expected kernel to be either a namedtuple, `Kernel`, or `np.ndarray`, got {type(b00)}.')


def _check_diag_batch(res):
  d = res.asdict() if isinstance(res, Kernel) else (res._asdict() if _is_nt(res) else {})
  if 'diagonal_batch' in d and not d['diagonal_batch']:
    raise NotImplementedError('Batching not implemented for `diagonal_batch == False`.')


def _native_matrix_output(kernel_fn, x1, get) -> bool:
  """True when `kernel_fn` is ours and the result is a set of [n1, n2] matrices, so a whole
  slab can be tiled inside one `ntk_gram_host` call."""
  spec = getattr(kernel_fn, '_spec', None)
  if spec is None or not isinstance(x1, np.ndarray) or get is None:
    return False
  names = (get,) if isinstance(get, str) else tuple(get)
  if not all(n.lower() in ('nngp', 'ntk') for n in names):
    return False
  from . import stax
  try:
    shape = stax._out_shape(spec, tuple(x1.shape))
  except Exception:
    return False
  return not isinstance(shape, list) and len(shape) == 2


def _serial_blocks(kernel_fn, x1, x2, n1_bs, n2_bs, args, kwargs):
  n1 = x1.shape[0]
  x2_is_none = x2 is None
  x2e = x1 if x2_is_none else x2
  n2 = x2e.shape[0]
  blocks = []
  for r in range(0, n1, n1_bs):
    row = []
    for c in range(0, n2, n2_bs):
      res = kernel_fn(x1[r:r + n1_bs], x2e[c:c + n2_bs], *args, **kwargs)
      _check_diag_batch(res)
      row.append(res)
    blocks.append(row)
  return _stitch(blocks, x2_is_none, n1, n2)


def _serial_kernel_blocks(kernel_fn, k: Kernel, n1_bs, n2_bs, args, kwargs):
  """`_src/batching.py:424-486`: Kernel-in batching via `Kernel.slice`."""
  n1 = k.cov1.shape[0]
  cov2_is_none = k.cov2 is None
  n2 = n1 if cov2_is_none else k.cov2.shape[0]
  blocks = []
  for r in range(0, n1, n1_bs):
    row = []
    for c in range(0, n2, n2_bs):
      res = kernel_fn(k.slice(slice(r, r + n1_bs), slice(c, c + n2_bs)), *args, **kwargs)
      _check_diag_batch(res)
      row.append(res)
    blocks.append(row)
  return _stitch(blocks, cov2_is_none, n1, n2)


def _symmetric_multi_device(kernel_fn, x, get, D):
  """`kernel_fn(x, None, get)` over D GPUs of this process: the triangular folded-cyclic schedule of
  `distributed.sym_schedule`, one host thread per device, mirrored on the host."""
  from . import distributed, stax
  names = (get,) if isinstance(get, str) else tuple(n.lower() for n in get)
  n = x.shape[0]
  sched = distributed.sym_schedule(n, D, distributed.sym_block_rows(n, D))
  out = {nm: None for nm in names}
  errs = []

  def work(dev):
    try:
      with _lib.device_scope(dev):
        for start, stop, r in sched:
          if r != dev:
            continue
          res = stax._sym_rows(kernel_fn, x, start, stop, names)
          for nm in names:
            out[nm][start:stop, start:] = res[nm]
    except BaseException as e:   # re-raised on the calling thread
      errs.append(e)

  from ._config import config
  for nm in names:
    out[nm] = np.empty((n, n), config.dtype)
  threads = [threading.Thread(target=work, args=(d,)) for d in range(D)]
  for t in threads:
    t.start()
  for t in threads:
    t.join()
  if errs:
    raise errs[0]
  for nm in names:
    m = out[nm]
    il = np.tril_indices(n, -1)
    m[il] = m.T[il]
  if isinstance(get, str):
    return out[names[0]]
  return stax._namedtuple(names)(*(out[nm] for nm in names))


def batch(kernel_fn: Callable, batch_size: int = 0, device_count: int = -1,
          store_on_device: bool = True) -> Callable:
  """Returns a function that computes a kernel in batches over all devices.

  Same signature and semantics as `neural_tangents.batch` (`_src/batching.py:76-132`):
  `x1.shape[0]` must be divisible by `device_count * batch_size` and `x2.shape[0]` by
  `batch_size` (`ValueError` otherwise).
  """
  del store_on_device  # results always land in host memory here
  if device_count == -1:
    n_dev = _lib.device_count()
    use_multidevice = n_dev > 1
    device_count_eff = n_dev if use_multidevice else 0
  else:
    use_multidevice = device_count > 0
    device_count_eff = device_count
  use_serial = bool(batch_size)
  D = device_count_eff if use_multidevice else 1

  def run_slab(dev, x1_slab, x2, get, args, kwargs, out, idx, n1_bs, n2_bs):
    try:
      with _lib.device_scope(dev):
        if isinstance(x1_slab, Kernel):
          if use_serial:
            out[idx] = _serial_kernel_blocks(kernel_fn, x1_slab, n1_bs, n2_bs, args, kwargs)
          else:
            out[idx] = kernel_fn(x1_slab, *args, **kwargs)
        elif not use_serial or _native_matrix_output(kernel_fn, x1_slab, get):
          out[idx] = kernel_fn(x1_slab, x2, *args, **kwargs)
        else:
          out[idx] = _serial_blocks(kernel_fn, x1_slab, x2, n1_bs, n2_bs, args, kwargs)
    except BaseException as e:  # re-raised on the calling thread
      out[idx] = e

  def batched_kernel_fn(x1_or_kernel, x2=None, *args, **kwargs):
    get = kwargs.get('get', args[0] if args else None)
    is_kernel = isinstance(x1_or_kernel, Kernel)
    if is_kernel:
      if x2 is not None:
        raise ValueError('x2 must be None when the first argument is a Kernel')
      n1 = x1_or_kernel.cov1.shape[0]
      n2 = n1 if x1_or_kernel.cov2 is None else x1_or_kernel.cov2.shape[0]
    elif isinstance(x1_or_kernel, np.ndarray):
      n1 = x1_or_kernel.shape[0]
      n2 = n1 if x2 is None else x2.shape[0]
    else:
      raise NotImplementedError(f'unsupported input type {type(x1_or_kernel)}')

    n1_bs = n2_bs = 0
    if use_serial:
      _, n1_bs_total, _, n2_bs = _get_n_batches_and_batch_sizes(n1, n2, batch_size, D)
      n1_bs = n1_bs_total // D
    if D <= 1:
      out = [None]
      run_slab(None, x1_or_kernel, x2, get, args, kwargs, out, 0, n1_bs, n2_bs)
      if isinstance(out[0], BaseException):
        raise out[0]
      return out[0]

    n_per_dev, d_eff = _get_n_per_device(n1, D)
    # x2 (or x1 itself when x2 is None) is shared by every device; rows are split in slabs.
    x2_is_none = x2 is None and not is_kernel
    if x2_is_none and d_eff > 1 and _native_matrix_output(kernel_fn, x1_or_kernel, get) and not args[1:] and \
        not (set(kwargs) - {'get'}):
      # K(x, x) on several GPUs: only the upper triangle is computed (the reference's TODO,
      # `_src/batching.py:370`), row blocks dealt to the devices in folded cyclic order for balance
      return _symmetric_multi_device(kernel_fn, x1_or_kernel, get, d_eff)
    x2_full = x1_or_kernel if x2_is_none else x2
    outs = [None] * d_eff
    threads = []
    for d in range(d_eff):
      sl = slice(d * n_per_dev, (d + 1) * n_per_dev)
      if is_kernel:
        slab = x1_or_kernel.slice(sl, slice(0, n2))
        a = (d, slab, None, get, args, kwargs, outs, d, n1_bs, n2_bs)
      else:
        a = (d, x1_or_kernel[sl], x2_full, get, args, kwargs, outs, d, n1_bs, n2_bs)
      t = threading.Thread(target=run_slab, args=a)
      t.start()
      threads.append(t)
    for t in threads:
      t.join()
    for o in outs:
      if isinstance(o, BaseException):
        raise o
    cov2_is_none = x1_or_kernel.cov2 is None if is_kernel else x2_is_none
    return _stitch([[o] for o in outs], cov2_is_none, n1, n2)

  batched_kernel_fn.input_req = getattr(kernel_fn, 'input_req', {})
  if hasattr(kernel_fn, '_spec'):
    batched_kernel_fn._spec = kernel_fn._spec
  batched_kernel_fn.device_count = device_count_eff
  return batched_kernel_fn


# ---- restartable on-disk Gram ---------------------------------------------------
def gram_to_disk(kernel_fn: Callable, x1: np.ndarray, x2, get, out_dir: str, block_rows: int = 512):
  """Computes `kernel_fn(x1, x2, get)` slab by slab ([block_rows, n2] row slabs) into `out_dir` and can be
  restarted: a slab whose file already exists is not recomputed.  Returns read-only `np.memmap`s of the
  assembled `[n1, n2]` matrices (a single one for a `str` `get`, else a namedtuple over `get`).

  The reference keeps every block in host memory until the final `jnp.stack` (`_src/batching.py:353-357,
  370` -- twice the peak memory and nothing survives an interruption); for Grams that outgrow the host
  (N >= 1e5) the slabs go to disk instead.  Files: `<name>.slab<row0>.npy` per finished slab (written
  atomically), `<name>.npy` for the assembled matrix, `manifest.json` with the shapes.
  """
  import collections
  import json
  import os
  names = (get,) if isinstance(get, str) else tuple(get)
  if not names or not all(isinstance(n, str) for n in names):
    raise ValueError('`get` must name the matrices to compute, e.g. ("nngp", "ntk").')
  if block_rows <= 0:
    raise ValueError('block_rows must be positive.')
  os.makedirs(out_dir, exist_ok=True)
  import hashlib
  from ._config import config
  n1 = x1.shape[0]
  n2 = n1 if x2 is None else x2.shape[0]
  symmetric = x2 is None

  def digest(a):
    return None if a is None else hashlib.sha1(np.ascontiguousarray(a).view(np.uint8)).hexdigest()
  spec = getattr(kernel_fn, '_spec', None)
  manifest = {'n1': int(n1), 'n2': int(n2), 'get': list(names), 'block_rows': int(block_rows),
              'symmetric': bool(symmetric), 'dtype': np.dtype(config.dtype).name if spec is not None else None,
              'x1_sha1': digest(x1), 'x2_sha1': digest(x2),
              'network_sha1': None if spec is None else hashlib.sha1(repr(spec).encode()).hexdigest()}
  mpath = os.path.join(out_dir, 'manifest.json')
  if os.path.exists(mpath):
    old = json.load(open(mpath))
    if any(old.get(k) != v for k, v in manifest.items()):
      diff = sorted(k for k, v in manifest.items() if old.get(k) != v)
      raise ValueError(f'{out_dir} holds a different computation (differs in {diff}); refusing to mix slabs.')
  else:
    with open(mpath, 'w') as f:
      json.dump(manifest, f)
  native_sym = symmetric and _native_matrix_output(kernel_fn, x1, get)

  def slab_path(name, r0):
    return os.path.join(out_dir, f'{name}.slab{r0:09d}.npy')

  # x2 = None: a slab holds only the columns [r0, n2) of its rows (upper trapezoid; the rest is mirrored at
  # assembly time), i.e. half the work of the reference's full square (`_src/batching.py:370`).
  for r0 in range(0, n1, block_rows):
    r1 = min(n1, r0 + block_rows)
    if all(os.path.exists(slab_path(n, r0)) for n in names):
      continue                                                  # finished in an earlier run
    c0 = r0 if symmetric else 0
    if native_sym:
      from . import stax
      d = stax._sym_rows(kernel_fn, x1, r0, r1, names)
      vals = [d[n] for n in names]
    else:
      res = kernel_fn(x1[r0:r1], x1[c0:] if symmetric else x2, names if len(names) > 1 else names[0])
      vals = [res] if len(names) == 1 else [getattr(res, n) for n in names]
    for n, v in zip(names, vals):
      v = np.asarray(v)
      if v.shape != (r1 - r0, n2 - c0):
        raise ValueError(f'`{n}` of rows [{r0}, {r1}) has shape {v.shape}, expected {(r1 - r0, n2 - c0)}.')
      tmp = slab_path(n, r0) + '.tmp.npy'
      np.save(tmp, v)
      os.replace(tmp, slab_path(n, r0))                         # atomic: a slab file is always complete

  outs = []
  for n in names:
    first = np.load(slab_path(n, 0), mmap_mode='r')
    full_path = os.path.join(out_dir, f'{n}.npy')
    full = np.lib.format.open_memmap(full_path, mode='w+', dtype=first.dtype, shape=(n1, n2))
    for r0 in range(0, n1, block_rows):
      sl = np.load(slab_path(n, r0), mmap_mode='r')
      rows = sl.shape[0]
      if symmetric:
        iu = np.triu_indices(rows)                              # entries (i, j >= i) of the leading square are valid
        sq = np.array(sl[:, :rows])
        sq.T[iu] = sq[iu]                                       # mirror inside the diagonal block
        full[r0:r0 + rows, r0:r0 + rows] = sq
        full[r0:r0 + rows, r0 + rows:] = sl[:, rows:]
        full[r0 + rows:, r0:r0 + rows] = sl[:, rows:].T         # mirror below the diagonal block
      else:
        full[r0:r0 + rows] = sl
    full.flush()
    del full
    outs.append(np.load(full_path, mmap_mode='r'))
  if isinstance(get, str):
    return outs[0]
  return collections.namedtuple('AnalyticKernel', names)(*outs)
