"""World-size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: shape broadcast, input broadcast,
contiguous row partition (x2 given), folded-cyclic triangular schedule (x2=None), slab all-gather and the
symmetric assembly.  The transport / per-block computation is a stand-in backend defined here (gloo + the
NumPy oracle); the schedule, layout and assembly logic under test is the product's `distributed.gram`.
The NCCL / CUDA backend is covered by the -m gpu tests (tests/test_gpu_parity.py, tests/test_multi_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


class GlooBackend:
  """`distributed.DeviceBackend` stand-in: arrays are NumPy, collectives are torch.distributed/gloo, a block is
  computed by the float64 oracle.  Entries the schedule declares unneeded (below the diagonal under
  `upper=True`) are poisoned with NaN, so an assembly that read them would fail the comparison."""

  def __init__(self, rank, world, spec):
    self.rank, self.world, self.spec = rank, world, spec
    self.blocks = []

  def bcast_meta(self, values, src):
    import torch
    import torch.distributed as dist
    t = torch.zeros(16, dtype=torch.int64)
    if self.rank == src:
      t[:len(values)] = torch.tensor(values, dtype=torch.int64)
    dist.broadcast(t, src=src)
    return [int(v) for v in t]

  def put(self, x, shape, dtype, src):
    return np.array(x, dtype) if self.rank == src else np.zeros(shape, dtype)

  def broadcast(self, d, src):
    import torch
    import torch.distributed as dist
    dist.broadcast(torch.from_numpy(d), src=src)     # in place (shares memory with the NumPy array)

  def alloc(self, shape, dtype):
    return np.full(shape, np.nan, dtype)

  def all_gather(self, slab):
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(slab))
    outs = [torch.empty_like(t) for _ in range(self.world)]
    dist.all_gather(outs, t)
    return torch.cat(outs, 0).numpy()

  def download(self, d):
    return np.array(d)

  def resolve(self, kernel_fn, x_shape):
    return kernel_fn

  def gram_block(self, plan, x1, r0, r1, x2, c0, outs, out_row, out_col, upper):
    from oracle import ntk_oracle as O
    self.blocks.append((r0, r1, c0, upper))
    k, t = O.kernel_fn(self.spec, x1[r0:r1], x2[c0:], ('nngp', 'ntk'), dtype=np.float64)
    if upper:
      assert x1 is x2 and c0 == r0
      low = np.tril_indices(r1 - r0, -1)
      k[low] = np.nan
      t[low] = np.nan
    for nm, v in (('nngp', k), ('ntk', t)):
      if nm in outs:
        outs[nm][out_row:out_row + (r1 - r0), out_col:out_col + v.shape[1]] = v

  def sym_assemble(self, slabs, row_of, n):
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    a, b = np.minimum(i, j), np.maximum(i, j)
    return slabs[row_of[a], b]

  def free(self, d):
    pass

  def synchronize(self):
    pass


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  import torch.distributed as dist
  import neural_tangents_b200 as nt
  from neural_tangents_b200 import distributed as D
  from oracle import ntk_oracle as O
  import cases
  nt.config.update('enable_x64', True)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    spec = cases.fcn(2, 1.5, 0.1)
    be = GlooBackend(rank, world, spec)
    rng = np.random.default_rng(0)
    n1 = 4 * world
    x1 = rng.standard_normal((n1, 12))
    x2 = rng.standard_normal((6, 12))
    ref = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'), dtype=np.float64)
    # inputs only on rank 0: the shape and data broadcasts must deliver them
    out = D.gram(spec, x1 if rank == 0 else None, x2 if rank == 0 else None, ('nngp', 'ntk'), backend=be)
    ok = np.allclose(out.nngp, ref[0], rtol=1e-12) and np.allclose(out.ntk, ref[1], rtol=1e-12)
    ok = ok and out.nngp.shape == (n1, 6)
    # contiguous row slab of this rank
    slab = D.gram(spec, x1, x2, 'ntk', backend=be, gather=False)
    lo, hi = D.row_partition(n1, world, rank)
    ok = ok and np.allclose(slab, ref[1][lo:hi], rtol=1e-12)
    # x2 = None: triangular folded-cyclic schedule; a ragged last block and several block heights
    sref = O.kernel_fn(spec, x1, None, ('nngp', 'ntk'), dtype=np.float64)
    for block in (None, 1, 3, n1):
      be.blocks.clear()
      sym = D.gram(spec, x1 if rank == 0 else None, None, ('nngp', 'ntk'), backend=be, block_rows=block)
      # exact-duplicate pairs (the diagonal) sit on the sqrt singularity of the Relu NTK:
      # the oracle itself reproduces them only to ~sqrt(eps) when the BLAS blocking changes with the block height
      off = ~np.eye(n1, dtype=bool)
      ok = ok and np.allclose(sym.nngp, sref[0], rtol=1e-12) and np.allclose(sym.ntk[off], sref[1][off], rtol=1e-12)
      ok = ok and np.allclose(np.diag(sym.ntk), np.diag(sref[1]), rtol=1e-6)
      ok = ok and not np.isnan(sym.ntk).any()
      ok = ok and all(up and c0 == r0 for r0, _, c0, up in be.blocks)       # only upper trapezoids were computed
      sched = D.sym_schedule(n1, world, block or D.sym_block_rows(n1, world))
      mine = sorted((s, e) for s, e, r in sched if r == rank)
      ok = ok and mine == sorted((r0, r1) for r0, r1, _, _ in be.blocks)
    try:
      D.gram(spec, x1[:n1 - 1], x2, ('nngp', 'ntk'), backend=be)
      ok = False
    except ValueError:
      pass
    q.put((rank, bool(ok)))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_partition_broadcast_gather_triangular_gloo(world):
  import torch.multiprocessing as mp
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(r, True) for r in range(world)]


def test_row_partition_arithmetic():
  from neural_tangents_b200 import distributed as D
  assert D.row_partition(16, 4, 1) == (4, 8)
  with pytest.raises(ValueError, match='must divide number of physical devices'):
    D.row_partition(10, 4, 0)


@pytest.mark.parametrize('n,world', [(10000, 8), (4096, 8), (4096, 2), (1000, 4), (37, 3), (5, 8)])
def test_triangular_schedule_covers_and_balances(n, world):
  """Every row is owned exactly once; the folded cyclic deal keeps the busiest rank within a few percent of
  the mean when n is large against the block height."""
  from neural_tangents_b200 import distributed as D
  block = D.sym_block_rows(n, world)
  sched = D.sym_schedule(n, world, block)
  assert sched[0][0] == 0 and sched[-1][1] == n
  assert all(a[1] == b[0] for a, b in zip(sched, sched[1:]))
  rows_pad, row_of, local = D.sym_layout(sched, world)
  assert len(set(row_of.tolist())) == n and row_of.max() < world * rows_pad
  for r in range(world):
    for start, stop, l0 in local[r]:
      assert (row_of[start:stop] == r * rows_pad + l0 + np.arange(stop - start)).all()
  work = D.sym_work(sched, world)
  assert sum(work) == n * (n + 1) // 2
  if n >= 1000:
    assert max(work) <= 1.02 * (sum(work) / world)
