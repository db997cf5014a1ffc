"""Benchmark of the analytic NNGP/NTK hot path (BASELINE.json: Myrtle-10 kernel entries/s).

  python bench.py --gpus N --steps K --warmup W            # ours (one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path

A *step* is one Gram block: every rank computes a [b1, b2] block of kernel entries of the
Myrtle-10 (32x32x3) NNGP+NTK Gram matrix from synthetic N(0,1) inputs (the nt.batch tiling
of the 10000x10000 configuration; entries/s does not depend on which block is computed).
Rows are partitioned across ranks with no data-path collective ("weak" scaling: each rank
owns its own slab of x1 rows, x2 is replicated).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

WORKLOADS = {
    # name: (myrtle depth, algorithmic elements/entry  )
    'myrtle5': (5, 4 * 32**4 + 4 * 16**4 + 4 * 8**4),
    'myrtle7': (7, 4 * 32**4 + 8 * 16**4 + 8 * 8**4),
    'myrtle10': (10, 8 * 32**4 + 12 * 16**4 + 12 * 8**4),
}


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    return json.load(open(path)), 'measured'
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


class ClockSampler(threading.Thread):
  """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
  Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.stop_flag = index, [], threading.Event()

  def run(self):
    while not self.stop_flag.is_set():
      try:
        out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                              '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(',')]
        if len(parts) >= 6:
          self.samples.append(parts)
      except Exception:
        pass
      self.stop_flag.wait(0.2)

  def summary(self):
    self.stop_flag.set()
    self.join(timeout=6)
    if not self.samples:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    sm = sorted(float(s[0]) for s in self.samples)
    reasons = set()
    for s in self.samples:
      for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[2:6]):
        if v.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]),
            'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_port_entries_per_s(depth, n_pairs_side, dtype=np.float64):
  """Times the NumPy oracle (port of the reference path) on a bounded sample."""
  from oracle import ntk_oracle as O
  import cases
  spec = cases.myrtle(depth)
  x1 = np.random.default_rng(0).standard_normal((n_pairs_side, 32, 32, 3)).astype(np.float32)
  x2 = np.random.default_rng(1).standard_normal((n_pairs_side, 32, 32, 3)).astype(np.float32)
  t0 = time.perf_counter()
  O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'), dtype=dtype)
  dt = time.perf_counter() - t0
  return n_pairs_side * n_pairs_side / dt, dt


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  depth, _ = WORKLOADS[args.workload]
  cores = len(os.sched_getaffinity(0))
  side = args.ref_side
  for _ in range(max(args.warmup, 0) and 1):
    cpu_port_entries_per_s(depth, 1)
  vals, t_tot = [], 0.0
  for _ in range(args.steps):
    v, dt = cpu_port_entries_per_s(depth, side)
    vals.append(v)
    t_tot += dt
  value = side * side * len(vals) / t_tot
  line = {
      'impl': 'reference', 'metric': 'kernel_entries_per_sec', 'value': value, 'unit': 'entries/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * t_tot / len(vals), 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': f'{args.workload}_32x32x3_nngp+ntk', 'block': [side, side]},
      'cpu_baseline': {'value': value, 'unit': 'entries/s', 'cores': cores, 'kind': 'port',
                       'sample': f'{side}x{side} pairs per step, NumPy float64 restatement of the '
                                 'reference path (reference needs JAX, not installable here)'},
      'e2e': {'value': value, 'unit': 'entries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


def run_ours(args):
  import torch
  import torch.distributed as dist
  import __graft_entry__ as g
  g.build()
  import neural_tangents_b200 as nt
  from neural_tangents_b200 import _lib, stax
  import cases

  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  nt.config.update('device', local)
  x64 = args.dtype == 'f64'
  nt.config.update('enable_x64', x64)
  np_dt = np.float64 if x64 else np.float32
  t_dt = torch.float64 if x64 else torch.float32
  sz = 8 if x64 else 4

  depth, elems = WORKLOADS[args.workload]
  _, _, kernel_fn = cases.build(cases.myrtle(depth), stax)
  low = stax._lowered(stax._strip(kernel_fn._spec), False, False, True)
  b1, b2 = args.block
  # synthetic inputs: every rank owns its own x1 row slab, x2 is shared
  x1_h = np.random.default_rng(100 + rank).standard_normal((b1, 32, 32, 3)).astype(np_dt)
  x2_h = np.random.default_rng(1).standard_normal((b2, 32, 32, 3)).astype(np_dt)
  ctx = _lib.get_context(local)
  stream = torch.cuda.ExternalStream(ctx.stream, device=local)
  x1_d = torch.from_numpy(x1_h).cuda(local)
  x2_d = torch.from_numpy(x2_h).cuda(local)
  nngp_d = torch.empty((b1, b2), dtype=t_dt, device=f'cuda:{local}')
  ntk_d = torch.empty((b1, b2), dtype=t_dt, device=f'cuda:{local}')
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=f'cuda:{local}')  # > 126 MB L2
  flags = _lib.FLAG_NO_FUSION if args.no_fusion else 0

  def step_device():
    _lib.gram_device(ctx, low.program, np_dt, x1_d.data_ptr(), b1, x2_d.data_ptr(), b2, 32, 32, 3,
                     flags, nngp_d.data_ptr(), ntk_d.data_ptr(), b2)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  torch.cuda.synchronize()
  for _ in range(args.warmup):
    step_device()
  ctx.synchronize()
  launches0 = ctx.launch_count
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  barrier()
  evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
         for _ in range(args.steps)]
  for s, e in evs:
    flush.zero_()                      # L2 flush between timed iterations (untimed)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
      s.record(stream)
      step_device()
      e.record(stream)
  barrier()
  ms_dev = sum(s.elapsed_time(e) for s, e in evs)
  launches = ctx.launch_count - launches0

  # end-to-end through the public API: host buffers in, host results out (nt.batch)
  batched = nt.batch(kernel_fn, batch_size=args.e2e_batch, device_count=0)
  batched(x1_h[:args.e2e_batch], x2_h[:args.e2e_batch], ('nngp', 'ntk'))
  barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    res = batched(x1_h, x2_h, ('nngp', 'ntk'))
  torch.cuda.synchronize()
  e2e_s = time.perf_counter() - t0
  barrier()
  clocks = sampler.summary() if rank == 0 else None

  t = torch.tensor([ms_dev, e2e_s * 1e3], dtype=torch.float64, device=f'cuda:{local}')
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_dev_max, e2e_ms_max = t.tolist()
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  entries_per_step = b1 * b2 * world
  value = entries_per_step * args.steps / (ms_dev_max * 1e-3)
  e2e_value = entries_per_step * args.steps / (e2e_ms_max * 1e-3)
  pk, pk_kind = peaks()
  bytes_per_entry = elems * sz
  achieved = (b1 * b2 * args.steps * bytes_per_entry) / (ms_dev / 1e3) / 1e9   # rank 0, GB/s
  # sanity: result must agree with the oracle-checked path (cheap spot check on 1 entry is in smoke())
  assert np.isfinite(res.nngp).all() and np.isfinite(res.ntk).all()
  cpu_v, cpu_dt = cpu_port_entries_per_s(depth, args.ref_side) if world == 1 else (None, None)
  line = {
      'metric': 'kernel_entries_per_sec', 'value': value, 'unit': 'entries/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev_max / args.steps,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype,
      'data': 'synthetic',
      'config': {'workload': f'{args.workload}_32x32x3_nngp+ntk', 'block_per_gpu': [b1, b2],
                 'parallelism': f'row-partition x{world}', 'l2': 'flushed between timed steps',
                 'fusion': not args.no_fusion},
      'e2e': {'value': e2e_value, 'unit': 'entries/s',
              'h2d_bytes_per_step': int(x1_h.nbytes + x2_h.nbytes) * world,
              'd2h_bytes_per_step': int(2 * b1 * b2 * sz) * world},
      'gpu_launches': int(launches),
      'clocks': clocks,
      'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                   'frac': achieved / pk['hbm_gbs'], 'traffic': None, 'peak_kind': pk_kind,
                   'kernel': 'whole layer pipeline (algorithmic bytes/entry x entries / step time)'},
  }
  if cpu_v is not None:
    line['cpu_baseline'] = {'value': cpu_v, 'unit': 'entries/s', 'cores': len(os.sched_getaffinity(0)),
                            'kind': 'port',
                            'sample': f'{args.ref_side}x{args.ref_side} pairs, NumPy float64 oracle, {cpu_dt:.1f}s'}
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='myrtle10', choices=sorted(WORKLOADS))
  ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
  ap.add_argument('--block', type=int, nargs=2, default=[64, 64])
  ap.add_argument('--e2e-batch', type=int, default=32)
  ap.add_argument('--ref-side', type=int, default=4)
  ap.add_argument('--no-fusion', action='store_true')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
