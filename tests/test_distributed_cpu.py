"""World-size-2 gloo test of the multi-GPU host logic (row partition, x2 broadcast, slab
all-gather) on CPU.  The kernel_fn here is a NumPy stand-in: the partition/collective
plumbing is what is under test (the CUDA path needs a GPU and is covered by -m gpu tests)."""
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


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  import collections
  import torch.distributed as dist
  from neural_tangents_b200 import distributed as D
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    AK = collections.namedtuple('AnalyticKernel', ('nngp', 'ntk'))

    def kernel_fn(x1, x2=None, get=None):
      x2 = x1 if x2 is None else x2
      k = x1.reshape(len(x1), -1) @ x2.reshape(len(x2), -1).T
      return AK(k, 2.0 * k)

    rng = np.random.default_rng(0)
    x1 = rng.standard_normal((8, 4, 4, 3)).astype(np.float32)
    x2 = rng.standard_normal((6, 4, 4, 3)).astype(np.float32)
    # inputs only on rank 0: broadcast must deliver them
    out = D.gram(kernel_fn, x1 if rank == 0 else None, x2 if rank == 0 else None, ('nngp', 'ntk'))
    ref = kernel_fn(x1, x2)
    ok = np.allclose(out.nngp, ref.nngp) and np.allclose(out.ntk, ref.ntk) and out.nngp.shape == (8, 6)
    sym = D.gram(kernel_fn, x1 if rank == 0 else None, None, ('nngp', 'ntk'))
    ok = ok and np.allclose(sym.nngp, kernel_fn(x1).nngp)
    slab = D.gram(kernel_fn, x1, x2, ('nngp', 'ntk'), gather=False)
    lo, hi = D.row_partition(8, world, rank)
    ok = ok and np.allclose(slab.nngp, ref.nngp[lo:hi])
    try:
      D.gram(kernel_fn, x1[:7], x2, ('nngp', 'ntk'))
      ok = False
    except ValueError:
      pass
    q.put((rank, bool(ok)))
  finally:
    dist.destroy_process_group()


def test_row_partition_broadcast_gather_gloo_world2():
  import torch.multiprocessing as mp
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(0, True), (1, True)]


def test_row_partition_arithmetic():
  from neural_tangents_b200 import distributed as D
  assert D.row_partition(16, 4, 1) == (4, 8)
  with pytest.raises(ValueError, match='must divide number of physical devices'):
    D.row_partition(10, 4, 0)
