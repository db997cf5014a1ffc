"""Multi-process Gram tiling: one rank per GPU, NCCL over NVLink through libntk_b200.so.

The reference's device parallelism is a single-process `jax.pmap` over x1 rows with x2 replicated
(`_src/batching.py:505-644,731-780`; no collectives anywhere), and it computes the full
square even when `x2 is None` (its own TODO at `_src/batching.py:370`).  The B200 layout is one process per
GPU:

  * inputs live on rank `src`; they are uploaded once and broadcast device-to-device (`ntk_comm_broadcast`,
    NCCL over NVLink); every rank derives the per-sample variance maps of the columns it needs locally;
  * `x2` given: rank r owns the contiguous row slab `[r n1/W, (r+1) n1/W)` (the reference's partition);
  * `x2 is None`: only the upper triangle is computed.  Rows are cut into blocks that are dealt to the ranks
    in folded (boustrophedon) cyclic order -- block I has `n - start_I` columns of work, and pairing block
    I with block 2W-1-I equalises the ranks.  Each block is one `ntk_gram_device` call with
    `NTK_FLAG_UPPER_ONLY` on `x[start_I:stop_I]` against `x[start_I:]`;
  * result slabs are all-gathered on the device (`ntk_comm_all_gather`), the symmetric matrix is put together
    by `ntk_sym_assemble`, and only then copied to the host.  Entry (i, j) depends on x1[i] and x2[j] alone:
    there is no reduction collective and no exchange inside the computation.

No PyTorch: rendezvous is a 128-byte NCCL id passed through a file (single node) or given by the caller;
`torchrun` is only the process launcher (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT in the environment).
The transport and the per-block computation sit behind a small backend interface so that the schedule and
the assembly are tested on CPU hosts with a gloo stand-in (tests/test_distributed_cpu.py).
"""
import os
import time

import numpy as np

from . import _lib
from ._config import config


# ---------------------------------------------------------------------------------------------------
# schedules (pure host logic)
# ---------------------------------------------------------------------------------------------------
def row_partition(n1: int, world: int, rank: int):
  """Contiguous slab of rank `rank`; `n1` must divide evenly (`_src/batching.py:560-573`)."""
  per, ragged = divmod(n1, world)
  if per and ragged:
    raise ValueError(('Dataset size ({}) must divide number of '
                      'physical devices ({}).').format(n1, world))
  if not per:
    raise ValueError(f'Dataset size ({n1}) is smaller than the number of ranks ({world}).')
  return rank * per, (rank + 1) * per


def sym_block_rows(n: int, world: int, target: int = 128) -> int:
  """Row-block height for the triangular schedule: the number of blocks is a multiple of 2*world whenever
  n allows it (then the folded assignment balances the ranks exactly), blocks are at most `target` rows."""
  if world <= 1:
    return n
  groups = max(1, -(-n // (2 * world * target)))        # ceil
  return max(1, -(-n // (2 * world * groups)))


def sym_schedule(n: int, world: int, block: int):
  """[(start, stop, rank)] for every row block, dealt in folded cyclic order 0..W-1, W-1..0, 0.."""
  out = []
  for I, start in enumerate(range(0, n, block)):
    q = I % (2 * world)
    out.append((start, min(n, start + block), q if q < world else 2 * world - 1 - q))
  return out


def sym_layout(schedule, world: int):
  """Where every global row lives after the all-gather.

  Returns (rows_pad, row_of, local): each rank stores its blocks back to back in a `[rows_pad, n]` slab
  (rows_pad = the largest per-rank row count); `row_of[i] = rank * rows_pad + local_row`; `local[rank]` is the
  list of (start, stop, local_row0) of that rank's blocks."""
  n = schedule[-1][1]
  local = [[] for _ in range(world)]
  fill = [0] * world
  for start, stop, r in schedule:
    local[r].append((start, stop, fill[r]))
    fill[r] += stop - start
  rows_pad = max(max(fill), 1)
  row_of = np.empty(n, np.int32)
  for r in range(world):
    for start, stop, l0 in local[r]:
      row_of[start:stop] = r * rows_pad + l0 + np.arange(stop - start, dtype=np.int32)
  return rows_pad, row_of, local


def sym_work(schedule, world: int):
  """Pairs computed per rank (for load-balance checks)."""
  n = schedule[-1][1]
  w = [0] * world
  for start, stop, r in schedule:
    rows = stop - start
    w[r] += rows * (n - start) - rows * (rows - 1) // 2
  return w


# ---------------------------------------------------------------------------------------------------
# rendezvous + communicator
# ---------------------------------------------------------------------------------------------------
_default = None


def _file_rendezvous(rank: int, world: int, tag: str, timeout: float = 300.0) -> bytes:
  """Rank 0 writes the NCCL id to a file every rank of this launch can name; the others poll for it."""
  path = os.path.join(os.environ.get('NTK_B200_RENDEZVOUS_DIR', '/tmp'), f'ntk_b200_nccl_{tag}.id')
  if rank == 0:
    uid = _lib.Comm.unique_id()
    tmp = f'{path}.{os.getpid()}.tmp'
    with open(tmp, 'wb') as f:
      f.write(uid)
    os.replace(tmp, path)                      # atomic: readers never see a partial id
    return uid
  t0 = time.time()
  while True:
    try:
      with open(path, 'rb') as f:
        uid = f.read()
      if len(uid) == _lib.COMM_ID_BYTES:
        return uid
    except FileNotFoundError:
      pass
    if time.time() - t0 > timeout:
      raise TimeoutError(f'rank {rank}: no NCCL id at {path} after {timeout:.0f} s')
    time.sleep(0.01)


def init(rank=None, world=None, local_rank=None, unique_id: bytes = None):
  """Creates (once) the process-wide communicator.  Defaults come from the launcher's environment
  (RANK, WORLD_SIZE, LOCAL_RANK as set by torchrun / mpirun wrappers); the NCCL id travels through a file
  named after MASTER_PORT and the launcher's pid unless `unique_id` is given."""
  global _default
  if _default is not None:
    return _default
  rank = int(os.environ.get('RANK', '0')) if rank is None else rank
  world = int(os.environ.get('WORLD_SIZE', '1')) if world is None else world
  local_rank = int(os.environ.get('LOCAL_RANK', str(rank))) if local_rank is None else local_rank
  ctx = _lib.get_context(local_rank)
  config.update('device', local_rank)
  tag = None
  if unique_id is None:
    tag = '{}_{}_{}'.format(os.environ.get('MASTER_PORT', '0'), os.environ.get('TORCHELASTIC_RUN_ID', 'run'),
                            os.getppid())
    unique_id = _file_rendezvous(rank, world, tag)
  _default = DeviceBackend(_lib.Comm(ctx, unique_id, rank, world))
  if rank == 0 and tag is not None:
    # every rank has read the id once the communicator exists (ncclCommInitRank is collective)
    try:
      os.remove(os.path.join(os.environ.get('NTK_B200_RENDEZVOUS_DIR', '/tmp'), f'ntk_b200_nccl_{tag}.id'))
    except Exception:
      pass
  return _default


def shutdown():
  global _default
  if _default is not None:
    _default.close()
    _default = None


# ---------------------------------------------------------------------------------------------------
# device backend (the product path)
# ---------------------------------------------------------------------------------------------------
class DeviceArray:
  """A row-major array in this rank's HBM, owned through the C-ABI (`ntk_device_malloc`)."""

  def __init__(self, ctx, shape, dtype):
    self.ctx, self.shape, self.dtype = ctx, tuple(int(s) for s in shape), np.dtype(dtype)
    self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
    self.ptr = ctx.malloc(self.nbytes)

  def row_ptr(self, row):
    stride = int(np.prod(self.shape[1:], dtype=np.int64)) * self.dtype.itemsize
    return self.ptr + row * stride

  def free(self):
    if self.ptr:
      self.ctx.free(self.ptr)
      self.ptr = 0

  def __del__(self):
    try:
      self.free()
    except Exception:
      pass


class DeviceBackend:
  """Transport (NCCL) + per-block computation (`ntk_gram_device`) of this rank."""

  def __init__(self, comm):
    self.comm, self.ctx = comm, comm.ctx
    self.rank, self.world = comm.rank, comm.world

  def close(self):
    self.comm.close()

  # -- transport
  def bcast_meta(self, values, src):
    """Broadcasts a short list of int64 (shapes) from `src`."""
    if self.world == 1:
      return list(values) + [0] * (16 - len(values))
    buf = np.zeros(16, np.int64)
    if self.rank == src:
      buf[:len(values)] = values
    d = DeviceArray(self.ctx, buf.shape, buf.dtype)
    self.ctx.h2d(d.ptr, buf)
    self.comm.broadcast(d.ptr, d.nbytes, src)
    self.ctx.d2h(buf, d.ptr)
    d.free()
    return [int(v) for v in buf]

  def put(self, x, shape, dtype, src):
    """Device array of `shape` on every rank; rank `src` fills it from the host array `x` (H2D)."""
    d = DeviceArray(self.ctx, shape, dtype)
    if self.rank == src:
      keep = self.ctx.h2d(d.ptr, np.ascontiguousarray(x, dtype))
      self.ctx.synchronize()       # the pageable source must outlive the copy
      del keep
    return d

  def broadcast(self, d, src):
    """`d` of rank `src` into every rank's `d`, device to device over NVLink."""
    self.comm.broadcast(d.ptr, d.nbytes, src)

  def all_gather_host(self, values):
    """Small host vector of float64 from every rank -> [world, len] (timings: max over ranks)."""
    v = np.ascontiguousarray(values, np.float64)
    s = DeviceArray(self.ctx, v.shape, v.dtype)
    r = DeviceArray(self.ctx, (self.world,) + v.shape, v.dtype)
    self.ctx.h2d(s.ptr, v)
    self.comm.all_gather(s.ptr, r.ptr, s.nbytes)
    out = self.ctx.d2h(np.empty(r.shape, np.float64), r.ptr)
    s.free()
    r.free()
    return out

  def barrier(self):
    self.all_gather_host([0.0])

  def alloc(self, shape, dtype):
    return DeviceArray(self.ctx, shape, dtype)

  def all_gather(self, slab):
    out = DeviceArray(self.ctx, (self.world * slab.shape[0],) + slab.shape[1:], slab.dtype)
    self.comm.all_gather(slab.ptr, out.ptr, slab.nbytes)
    return out

  def download(self, d):
    out = _lib.pinned_empty(d.shape, d.dtype)
    self.ctx.d2h(out, d.ptr)
    return out

  # -- computation
  def resolve(self, kernel_fn, x_shape):
    """kernel_fn -> (lowered program, H, W, C); only [n1, n2] matrix outputs shard this way."""
    from . import stax
    spec = getattr(kernel_fn, '_spec', None)
    if spec is None:
      raise TypeError('distributed.gram needs a kernel_fn built by neural_tangents_b200.stax')
    spatial = len(x_shape) == 4
    low = stax._lowered(stax._strip(spec), False, False, spatial)
    H, W = (x_shape[1], x_shape[2]) if spatial else (0, 0)
    oh, _, _ = low.program.output_shape(H, W)
    if oh > 0:
      raise NotImplementedError('distributed.gram computes [n1, n2] matrices; this network keeps spatial axes')
    return low, H, W, x_shape[-1]

  def gram_block(self, plan, x1, r0, r1, x2, c0, outs, out_row, out_col, upper):
    """outs[name][out_row : out_row + (r1-r0), out_col : out_col + (n2-c0)] = K(x1[r0:r1], x2[c0:])."""
    low, H, W, C = plan
    flags = (_lib.FLAG_UPPER_ONLY if upper else 0) | (_lib.FLAG_NO_FUSION if config.disable_fusion else 0)
    k, t = outs.get('nngp'), outs.get('ntk')
    scratch = None
    if k is None:                 # ntk only: the C entry always produces nngp as well
      scratch = k = DeviceArray(self.ctx, t.shape, t.dtype)
    ld = k.shape[1]
    off = lambda a: a.ptr + (out_row * ld + out_col) * a.dtype.itemsize
    n2 = x2.shape[0] - c0
    _lib.gram_device(self.ctx, low.program, k.dtype, x1.row_ptr(r0), r1 - r0, x2.row_ptr(c0), n2, H, W, C, flags,
                     off(k), off(t) if t is not None else None, ld)
    if scratch is not None:
      self.ctx.synchronize()
      scratch.free()

  def sym_assemble(self, slabs, row_of, n):
    ro = DeviceArray(self.ctx, row_of.shape, np.int32)
    keep = self.ctx.h2d(ro.ptr, row_of)
    out = DeviceArray(self.ctx, (n, n), slabs.dtype)
    _lib.sym_assemble(self.ctx, slabs.dtype, slabs.ptr, slabs.shape[1], ro.ptr, n, out.ptr, n)
    self.ctx.synchronize()         # `row_of` and the gathered slabs are released by the caller right after
    del keep
    ro.free()
    return out

  def free(self, d):
    d.free()

  def synchronize(self):
    self.ctx.synchronize()


# ---------------------------------------------------------------------------------------------------
# the distributed Gram
# ---------------------------------------------------------------------------------------------------
def _names(get):
  names = (get,) if isinstance(get, str) else tuple(get)
  names = tuple(n.lower() for n in names)
  if not names or not all(n in ('nngp', 'ntk') for n in names):
    raise NotImplementedError('distributed.gram gathers the [n1, n2] matrices "nngp" and / or "ntk"')
  return names


def _result(get, names, vals):
  if isinstance(get, str):
    return vals[0]
  import collections
  return collections.namedtuple('AnalyticKernel', names)(*vals)


def gram_resident(be, plan, d1, d2, names, src=0, gather=True, block_rows=None, broadcast=True):
  """The device-resident core: `d1` / `d2` (None: symmetric) are device arrays on every rank, filled on rank
  `src`.  Broadcast (NCCL), per-rank blocks (`ntk_gram_device`), all-gather, symmetric assembly.  Returns
  {name: device array}; everything is enqueued on the rank's context stream and the call returns without a
  host synchronisation of the results."""
  dt = d1.dtype
  n1 = d1.shape[0]
  if broadcast and be.world > 1:
    be.broadcast(d1, src)
    if d2 is not None:
      be.broadcast(d2, src)
  if d2 is None:
    n = n1
    block = block_rows or sym_block_rows(n, be.world)
    sched = sym_schedule(n, be.world, block)
    rows_pad, row_of, local = sym_layout(sched, be.world)
    slabs = {nm: be.alloc((rows_pad, n), dt) for nm in names}
    for start, stop, l0 in local[be.rank]:
      be.gram_block(plan, d1, start, stop, d1, start, slabs, l0, start, upper=True)
    if not gather:
      return slabs
    outs = {}
    for nm in names:
      g = be.all_gather(slabs[nm]) if be.world > 1 else slabs[nm]
      outs[nm] = be.sym_assemble(g, row_of, n)
      if g is not slabs[nm]:
        be.free(g)
      be.free(slabs[nm])
    return outs
  n2 = d2.shape[0]
  lo, hi = row_partition(n1, be.world, be.rank)
  slabs = {nm: be.alloc((hi - lo, n2), dt) for nm in names}
  be.gram_block(plan, d1, lo, hi, d2, 0, slabs, 0, 0, upper=False)
  if not gather or be.world == 1:
    return slabs
  outs = {}
  for nm in names:
    outs[nm] = be.all_gather(slabs[nm])
    be.free(slabs[nm])
  return outs


def gram(kernel_fn, x1, x2=None, get=('nngp', 'ntk'), backend=None, src=0, gather=True, to_host=True,
         block_rows=None):
  """`kernel_fn(x1, x2, get)` with the Gram rows partitioned over the ranks of `backend` (default: the
  process-wide NCCL communicator, `init()`).

  `x1` / `x2` need only be given on rank `src` (other ranks may pass None).  Returns on every rank the full
  `[n1, n2]` matrices (`gather=True`) or this rank's row slab (`gather=False`; for `x2 is None` the slab holds
  this rank's row blocks of the triangular schedule, upper entries only).  `to_host=False` leaves the results
  in HBM (`DeviceArray`s) for a consumer on the device.
  """
  be = backend if backend is not None else init()
  names = _names(get)
  dt = np.dtype(config.dtype)
  # shapes travel first: [ndim, *x1.shape, has_x2, n2]
  meta = None
  if be.rank == src:
    if x1 is None:
      raise ValueError(f'x1 must be given on rank {src}')
    meta = [x1.ndim] + list(x1.shape) + [0 if x2 is None else 1, 0 if x2 is None else x2.shape[0]]
  meta = be.bcast_meta(meta, src)
  nd = meta[0]
  shape1 = tuple(meta[1:1 + nd])
  has_x2, n2 = bool(meta[1 + nd]), meta[2 + nd]
  plan = be.resolve(kernel_fn, shape1)
  d1 = be.put(x1, shape1, dt, src)
  d2 = be.put(x2, (n2,) + shape1[1:], dt, src) if has_x2 else None
  try:
    outs = gram_resident(be, plan, d1, d2, names, src=src, gather=gather, block_rows=block_rows)
    be.synchronize()
  finally:
    be.free(d1)
    if d2 is not None:
      be.free(d2)
  if not to_host:
    return _result(get, names, [outs[nm] for nm in names])
  vals = []
  for nm in names:
    vals.append(be.download(outs[nm]))
    be.free(outs[nm])
  return _result(get, names, vals)
