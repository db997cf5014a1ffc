"""Multi-process Gram tiling: one rank per GPU over `torch.distributed`.

The reference's device parallelism is a single-process `jax.pmap` over x1 rows with x2
replicated (`_src/batching.py:505-644,731-780`; no collectives anywhere).  The
B200 layout is one process per GPU: rank r owns the contiguous row slab
`[r*n1/W, (r+1)*n1/W)` of the Gram matrix, x2 (and with it the per-sample variance maps,
which every rank derives locally from x2) is broadcast once over NVLink, and the
`[n1/W, n2]` result slabs are all-gathered.  Entry (i, j) depends only on x1[i] and x2[j],
so there is no reduction collective and no exchange inside the computation.

`torch.distributed` is plumbing only (NCCL broadcast / all_gather, or gloo on CPU hosts);
every kernel entry is computed by `kernel_fn`, i.e. by libntk_b200.so on this rank's GPU.
"""
import numpy as np


def _dist():
  import torch.distributed as dist
  return dist


def row_partition(n1: int, world: int, rank: int):
  """Contiguous slab of rank `rank`; `n1` must divide evenly (`_src/batching.py:560-573`)."""
  per, ragged = divmod(n1, world)
  if per and ragged:
    raise ValueError(('Dataset size ({}) must divide number of '
                      'physical devices ({}).').format(n1, world))
  if not per:
    raise ValueError(f'Dataset size ({n1}) is smaller than the number of ranks ({world}).')
  return rank * per, (rank + 1) * per


def _device_for(group):
  import torch
  dist = _dist()
  if dist.get_backend(group) == 'nccl':
    return torch.device('cuda', torch.cuda.current_device())
  return torch.device('cpu')


def broadcast_array(x, src=0, group=None):
  """Broadcasts a NumPy array (shape/dtype included) from `src` to every rank."""
  import torch
  dist = _dist()
  rank = dist.get_rank(group)
  meta = [None if x is None else (tuple(x.shape), str(x.dtype))] if rank == src else [None]
  dist.broadcast_object_list(meta, src=src, group=group)
  if meta[0] is None:
    return None
  shape, dtype = meta[0]
  dev = _device_for(group)
  if rank == src:
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
  else:
    t = torch.empty(shape, dtype=getattr(torch, dtype), device=dev)
  dist.broadcast(t, src=src, group=group)
  return t.cpu().numpy()


def all_gather_rows(slab: np.ndarray, group=None) -> np.ndarray:
  """Concatenates equally-shaped row slabs of every rank along axis 0."""
  import torch
  dist = _dist()
  world = dist.get_world_size(group)
  dev = _device_for(group)
  t = torch.from_numpy(np.ascontiguousarray(slab)).to(dev)
  out = [torch.empty_like(t) for _ in range(world)]
  dist.all_gather(out, t, group=group)
  return torch.cat(out, dim=0).cpu().numpy()


def gram(kernel_fn, x1, x2=None, get=('nngp', 'ntk'), group=None, src=0, gather=True):
  """`kernel_fn(x1, x2, get)` with the rows of x1 partitioned over the ranks of `group`.

  `x1` / `x2` need only be given on rank `src` (other ranks may pass None).  Returns, on every
  rank, the full `[n1, n2]` matrices (`gather=True`) or this rank's `[n1/W, n2]` slab.
  """
  dist = _dist()
  if not dist.is_initialized():
    raise RuntimeError('torch.distributed is not initialised')
  world, rank = dist.get_world_size(group), dist.get_rank(group)
  x1 = broadcast_array(x1, src, group)
  x2 = broadcast_array(x2, src, group)
  lo, hi = row_partition(x1.shape[0], world, rank)
  x2_eff = x1 if x2 is None else x2
  res = kernel_fn(x1[lo:hi], x2_eff, get)
  if not gather:
    return res
  if isinstance(res, np.ndarray):
    return all_gather_rows(res, group)
  if hasattr(res, '_fields'):
    return type(res)(*(all_gather_rows(np.asarray(v), group) for v in res))
  raise NotImplementedError('distributed.gram gathers arrays / AnalyticKernel tuples of [n1, n2] matrices')
