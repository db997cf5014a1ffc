"""Whole-job run of a BASELINE config: the full N x N Myrtle-10 NNGP+NTK Gram matrix of N synthetic
32x32x3 inputs (BASELINE.json configs[3]: N = 10000), rows partitioned over the ranks
(`neural_tangents_b200.distributed.gram`: x broadcast with NCCL, slabs all-gathered, no reduction).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29533 profiles/full_gram.py --size 10000 [--dtype f64]

Prints one JSON line on rank 0: wall time (barrier to barrier, max over ranks), entries/s, and
size-independent checks of the result: symmetry of K(x, x), agreement of randomly chosen entries with a
direct 1-GPU recomputation and with the NumPy float64 oracle, and a PSD check on a random principal
512 x 512 sub-matrix.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--size', type=int, default=10000)
  ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
  ap.add_argument('--depth', type=int, default=10)
  ap.add_argument('--oracle-entries', type=int, default=4)
  args = ap.parse_args()

  import torch
  import torch.distributed as dist
  import __graft_entry__ as g
  g.build()
  import neural_tangents_b200 as nt
  from neural_tangents_b200 import distributed, stax
  import cases

  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if 'MASTER_ADDR' not in os.environ:  # plain `python profiles/full_gram.py`: one rank
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT='29534', RANK='0', WORLD_SIZE='1')
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  else:
    dist.init_process_group('gloo')
  nt.config.update('device', local)
  nt.config.update('enable_x64', args.dtype == 'f64')
  spec = cases.myrtle(args.depth)
  _, _, kernel_fn = cases.build(spec, stax)

  n = args.size - args.size % world
  x = np.random.default_rng(0).standard_normal((n, 32, 32, 3)).astype(np.float32) if rank == 0 else None
  # warm-up (context, workspace, module load) on a tiny problem
  distributed.gram(kernel_fn, None if rank else x[:world * 2], None if rank else x[:4], ('nngp', 'ntk'))
  dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  # x2 = x explicitly: every rank computes its [n/W, n] rectangle (the reference's schedule, batching.py:370)
  res = distributed.gram(kernel_fn, x, None if rank else x, ('nngp', 'ntk'))
  torch.cuda.synchronize()
  dist.barrier()
  dt = time.perf_counter() - t0
  t = torch.tensor([dt], dtype=torch.float64, device='cuda' if world > 1 else 'cpu')
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  dt = float(t.item())
  if rank == 0:
    nngp, ntk = np.asarray(res.nngp), np.asarray(res.ntk)
    out = {'workload': f'myrtle{args.depth}_32x32x3_nngp+ntk full Gram', 'n': n, 'n_gpus': world, 'dtype': args.dtype,
           'wall_s': dt, 'entries_per_s': n * n / dt, 'shape': list(nngp.shape)}
    # symmetry of K(x, x) computed as a full rectangle by independent ranks
    out['max_rel_asymmetry'] = float(max(np.abs(nngp - nngp.T).max() / np.abs(nngp).max(),
                                         np.abs(ntk - ntk.T).max() / np.abs(ntk).max()))
    out['finite'] = bool(np.isfinite(nngp).all() and np.isfinite(ntk).all())
    rng = np.random.default_rng(7)
    # random entries against a direct recomputation of small blocks on this GPU
    ii, jj = rng.choice(n, 6, replace=False), rng.choice(n, 5, replace=False)
    blk = kernel_fn(x[ii], x[jj], ('nngp', 'ntk'))
    out['max_rel_diff_vs_direct_block'] = float(max(np.abs(blk.nngp / nngp[np.ix_(ii, jj)] - 1).max(),
                                                    np.abs(blk.ntk / ntk[np.ix_(ii, jj)] - 1).max()))
    # a few entries against the NumPy float64 oracle (CPU)
    if args.oracle_entries:
      from oracle import ntk_oracle as O
      io, jo = ii[:2], jj[:max(1, args.oracle_entries // 2)]
      ref = O.kernel_fn(spec, x[io], x[jo], ('nngp', 'ntk'))
      out['max_rel_err_vs_oracle'] = float(max(np.abs(nngp[np.ix_(io, jo)] / ref[0] - 1).max(),
                                               np.abs(ntk[np.ix_(io, jo)] / ref[1] - 1).max()))
    # PSD on a random principal sub-matrix
    m = min(512, n)
    sel = np.sort(rng.choice(n, m, replace=False))
    for name, mat in (('nngp', nngp), ('ntk', ntk)):
      sub = mat[np.ix_(sel, sel)].astype(np.float64)
      w = np.linalg.eigvalsh((sub + sub.T) / 2)
      out[f'min_eig_over_max_{name}'] = float(w.min() / w.max())
    print(json.dumps(out))
  dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
