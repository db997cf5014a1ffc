"""Benchmark of the analytic NNGP/NTK hot path (BASELINE.json: Myrtle-10 kernel entries/s).

  python bench.py --gpus N --steps K --warmup W            # ours (one rank per GPU; torchrun launches N > 1)
  python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path

A *step* is one Gram block per rank: a [b1, b2] block of NNGP+NTK kernel entries of the Myrtle-10 (32x32x3)
network on synthetic N(0,1) inputs, i.e. one block of the nt.batch tiling of the 10000x10000 configuration
(entries/s does not depend on which block).  Multi-GPU runs go through the PRODUCT path,
`neural_tangents_b200.distributed` (NCCL inside libntk_b200.so; no torch anywhere in this file): x1
(world x b1 rows) and x2 live on rank 0, are broadcast device-to-device, every rank computes its row slab
and the slabs are all-gathered, all inside the timed region ("weak" scaling, no reduction collective).

Besides the headline line the JSON carries
  `roofline`  the dominant kernel (first fused stage), timed live with CUDA events on its launch stream,
              plus `roofline.compute`: its real bound after cross-layer fusion (issue / FMA pipe / register file);
  `strong`    a fixed-size symmetric Gram (x2=None, triangular folded-cyclic schedule) over all N GPUs through
              `distributed.gram`, host arrays in and out -- shows tail imbalance and collective cost;
  `configs`   one measured line per BASELINE.json config family (N = 1 only);
  `cpu_baseline`  the NumPy float64 port of the reference path on the host cores (N = 1 only).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

E32, E16, E8 = 32**4, 16**4, 8**4
# name: (myrtle depth, algorithmic elements/entry for the whole net, for fused stage 0)  SURVEY §8d:
# one HBM round trip (read nngp+ntk, write nngp+ntk) per Conv+Relu(+AvgPool) layer, layer 1 write-only.
WORKLOADS = {
    'myrtle5': (5, 4 * E32 + 4 * E16 + 4 * E8, 4 * E32 + 2 * E16),
    'myrtle7': (7, 4 * E32 + 8 * E16 + 8 * E8, 4 * E32 + 2 * E16),
    'myrtle10': (10, 8 * E32 + 12 * E16 + 12 * E8, 8 * E32 + 2 * E16),
    # README.md:370-395: 21 x [Conv3x3 SAME + Relu] + GlobalAvgPool, the only network the reference
    # publishes GPU timings for (V100 fp32: 2.7001 ms / NTK entry = 370.4 entries/s).
    'readme21': (21, 80 * E32, 10 * E32),
    # README.md:192-222 WideResnet(block_size=4, k=1); SURVEY §8d: 46 E32 + 42 E16 + 38 E8
    'wrn': (0, 46 * E32 + 42 * E16 + 38 * E8, 0),
    # README.md:399-416: the same 21 layers with Flatten on top (no pooling => only the diagonal
    # column is needed; the reference's diagonal_spatial path).  V100 fp32: 0.0046510 ms / entry.
    'readme21_flatten': (-1, 80 * 32 * 32, 0),
    # Erf variants (BASELINE config 5 names Relu/Erf): same algorithmic traffic as their Relu twins
    'wrn_erf': (-2, 46 * E32 + 42 * E16 + 38 * E8, 0),
    'myrtle10_erf': (10, 8 * E32 + 12 * E16 + 12 * E8, 8 * E32 + 2 * E16),
    # BASELINE configs[0]: infinite FCN (Dense-Relu x3 + Dense), 784-d inputs (examples/infinite_fcn.py
    # deepened x3).  Input Gram on the tensor cores (tcgen05 3xTF32 / DMMA fp64); the
    # per-entry epilogue is O(1), so the "algorithmic traffic" is just x1, x2 and the two result matrices.
    'fcn': (-4, 0, 0),
}
FCN_DIM = 784
PUBLISHED = {('readme21', 'f32'): 1e3 / 2.7001, ('readme21', 'f64'): 1e3 / 6.2058,
             ('readme21_flatten', 'f32'): 1e3 / 0.0046510, ('readme21_flatten', 'f64'): 1e3 / 0.010822}
# Real bound of the dominant fp32 kernel (ncu capture named below): static context for
# `roofline.compute`, the live part is the clock count per element-layer.
RF_FLOOR_CLK = {'f32': 35.0}


def input_shape(name, n):
  return (n, FCN_DIM) if name == 'fcn' else (n, 32, 32, 3)


def input_dims(name):
  return (0, 0, FCN_DIM) if name == 'fcn' else (32, 32, 3)


def workload_spec(name):
  import cases
  if name == 'readme21':
    return ('serial', [cases.conv(W=1., b=None), cases.RELU] * 21 + [('gap',)])
  if name == 'readme21_flatten':
    return ('serial', [cases.conv(W=1., b=None), cases.RELU] * 21 + [('flatten',)])
  if name == 'fcn':
    return cases.fcn(3, 2., 0.05)
  if name == 'myrtle10_erf':
    spec = cases.myrtle(10)
    return ('serial', [('erf', 1., 1., 0.) if l == cases.RELU else l for l in spec[1]])
  if name in ('wrn', 'wrn_erf'):
    act = cases.RELU if name == 'wrn' else ('erf', 1., 1., 0.)

    def group(n, stride):
      return [cases.wrn_block(stride, True, act)] + [cases.wrn_block(1, False, act) for _ in range(n - 1)]
    # Conv defaults of the reference: W_std = 1, b_std = None (cases.wrn_block uses b = 0.1, which
    # exercises the bias path as well)
    return ('serial', [cases.conv(W=1., b=None)] + group(4, 1) + group(4, 2) + group(4, 2) +
            [cases.pool((8, 8), (1, 1)), ('flatten',), ('dense', 1., 0.)])
  return cases.myrtle(WORKLOADS[name][0])


def workload_label(name):
  return f'{name}_{FCN_DIM}d_nngp+ntk' if name == 'fcn' else f'{name}_32x32x3_nngp+ntk'


def stage_elements(depth, per_layer):
  """Algorithmic elements (K and T, read + written) per pair for every stage kernel launch of the
  cross-pair pass, in launch order (SURVEY §8d traffic model)."""
  if depth == 21:   # readme21: 21 layers at 32x32, chunks of 3 (or 1), GAP at the end
    n = 21 if per_layer else 7
    per = 1 if per_layer else 3
    return [(0 if i == 0 else 2 * E32) + (per - 1) * 4 * E32 + (0 if i == n - 1 else 2 * E32)
            for i in range(n)]
  f = {5: [2, 1, 1], 7: [2, 2, 2], 10: [3, 3, 3]}[depth]
  res = [E32, E16, E8]
  out = []
  for g, (n_layers, e) in enumerate(zip(f, res)):
    e_next = res[g + 1] if g + 1 < 3 else 0          # after the pool (last group: 2 scalars)
    chunks = [1] * n_layers if per_layer else [n_layers]
    done = 0
    for c in chunks:
      first = g == 0 and done == 0
      last = done + c == n_layers
      if per_layer:
        rd = 0 if first else 2 * e
        wr = 2 * (e_next if last else e)
        out.append(rd + wr)
      else:                                            # the whole group in one launch
        rd = 0 if first else 2 * e
        out.append(rd + (n_layers - 1) * 4 * e + 2 * e_next)
      done += c
  return out


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    return json.load(open(path)), 'measured'
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


def ncu_capture(workload, dtype):
  """The committed `ncu --set full` summary of the dominant kernel (profiles/ncu_stage0_summary.json):
  dram bytes per pair and the pipe / issue utilisation, each labelled with the capture it comes from."""
  path = os.path.join(ROOT, 'profiles', 'ncu_stage0_summary.json')
  try:
    return json.load(open(path)).get(f'{workload}_{dtype}')
  except Exception:
    return None


class ClockSampler(threading.Thread):
  """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
  Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap,power.draw')

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
      self.stop_flag.wait(0.1)

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
    out = {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]),
           'reasons': sorted(reasons), 'samples': len(sm)}
    try:
      out['power_w_max'] = max(float(s[6]) for s in self.samples)
    except Exception:
      pass
    return out


# ---------------------------------------------------------------------------------------
# CPU arm: the NumPy float64 oracle (port of the reference path) on all host cores
# ---------------------------------------------------------------------------------------
def _cpu_worker(job):
  name, seed, n_cols = job
  os.environ.setdefault('OMP_NUM_THREADS', '1')
  from oracle import ntk_oracle as O
  spec = workload_spec(name)
  if name == 'fcn':
    n_cols *= 250   # an FCN entry costs ~1e-4 of a Myrtle-10 entry on the CPU
  x1 = np.random.default_rng(1000 + seed).standard_normal(input_shape(name, 1 if name != 'fcn' else 250)).astype(np.float32)
  x2 = np.random.default_rng(1).standard_normal(input_shape(name, n_cols)).astype(np.float32)
  out = O.kernel_fn(spec, x1, x2, ('nngp', 'ntk'), dtype=np.float64)
  return float(out[0].sum() + out[1].sum()), int(out[0].size)


def cpu_port_step(pool, name, cores, n_cols):
  """One bounded CPU step: `cores` workers, each one x1 row against `n_cols` x2 columns."""
  t0 = time.perf_counter()
  res = pool.map(_cpu_worker, [(name, s, n_cols) for s in range(cores)])
  dt = time.perf_counter() - t0
  assert all(np.isfinite(r[0]) for r in res)
  return sum(r[1] for r in res), dt


def make_cpu_pool(cores):
  import multiprocessing as mp
  return mp.get_context('spawn').Pool(cores)


def make_config(args, world, b1, b2):
  """`config` of the JSON line: the workload both arms are quoted on (the reference arm runs a bounded sample of it,
  described in its `cpu_baseline.sample`)."""
  return {'workload': workload_label(args.workload), 'block_per_gpu': [b1, b2],
          'parallelism': (f'x1-row partition over {world} rank(s): x1 / x2 broadcast from rank 0 (NCCL), contiguous row '
                          'slabs, slabs all-gathered (no reduction collective)') if world > 1 else 'one rank',
          'l2': 'flushed (256 MiB memset) between timed steps',
          'fusion': ('per-layer stencil kernels' if args.per_layer else not args.no_fusion),
          'symmetric_x2_none': bool(args.symmetric)}


def block_of(args):
  b1, b2 = args.block
  if args.workload == 'fcn' and args.block == [256, 256]:
    b1 = b2 = 1000                                                  # BASELINE configs[0]: 1000 x 1000
  return b1, b2


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cores = len(os.sched_getaffinity(0))
  pool = make_cpu_pool(cores)
  try:
    for _ in range(min(args.warmup, 1)):
      cpu_port_step(pool, args.workload, cores, 1)
    entries, t_tot = 0, 0.0
    for _ in range(args.steps):
      n, dt = cpu_port_step(pool, args.workload, cores, args.ref_cols)
      entries += n
      t_tot += dt
  finally:
    pool.terminate()
  value = entries / t_tot
  sample = (f'{cores} worker processes x (1 x {args.ref_cols}) pairs per step, NumPy float64 '
            'restatement of the reference path (the reference itself needs JAX, not installable here)')
  line = {
      'impl': 'reference', 'metric': 'kernel_entries_per_sec', 'value': value, 'unit': 'entries/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * t_tot / args.steps, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': make_config(args, args.gpus, *block_of(args)),
      'cpu_baseline': {'value': value, 'unit': 'entries/s', 'cores': cores, 'kind': 'port',
                       'sample': sample},
      'e2e': {'value': value, 'unit': 'entries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  emit(line)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
class Bench:
  """One rank of the GPU arm: the library context, the NCCL backend (world > 1) and timing helpers."""

  def __init__(self, args):
    import __graft_entry__ as g
    g.build()
    import neural_tangents_b200 as nt
    from neural_tangents_b200 import _lib, distributed, stax
    import cases
    self.nt, self.lib, self.D, self.stax, self.cases = nt, _lib, distributed, stax, cases
    self.args = args
    self.rank = int(os.environ.get('RANK', '0'))
    self.world = int(os.environ.get('WORLD_SIZE', '1'))
    self.local = int(os.environ.get('LOCAL_RANK', '0'))
    nt.config.update('device', self.local)
    self.ctx = _lib.get_context(self.local)
    self.be = distributed.init() if self.world > 1 else None
    self.flush_bytes = 256 << 20                                   # > 126 MB L2
    self.flush = self.ctx.malloc(self.flush_bytes)

  def set_dtype(self, dtype):
    self.x64 = dtype == 'f64'
    self.nt.config.update('enable_x64', self.x64)
    self.np_dt = np.float64 if self.x64 else np.float32
    self.sz = 8 if self.x64 else 4

  def kernel(self, workload):
    _, _, kernel_fn = self.cases.build(workload_spec(workload), self.stax)
    low = self.stax._lowered(self.stax._strip(kernel_fn._spec), False, False, workload != 'fcn')
    return kernel_fn, low

  def barrier(self):
    self.ctx.synchronize()
    if self.be is not None:
      self.be.barrier()

  def max_over_ranks(self, values):
    if self.be is None:
      return list(values)
    return self.be.all_gather_host(values).max(axis=0).tolist()

  def time_steps(self, step, steps, warmup):
    """`warmup` untimed + `steps` timed calls of `step()` on the context stream: CUDA events around every
    step, L2 flushed (256 MiB memset) outside the events, barrier + synchronize on both sides.
    Returns this rank's summed device ms."""
    for _ in range(warmup):
      step()
    self.barrier()
    evs = [(self.lib.Event(), self.lib.Event()) for _ in range(steps)]
    for s, e in evs:
      self.ctx.memset(self.flush, 0, self.flush_bytes)
      s.record(self.ctx)
      step()
      e.record(self.ctx)
    self.barrier()
    return sum(s.elapsed_ms(e) for s, e in evs)


def measure_block(B, workload, dtype, b1, b2, steps, warmup, flags=0, profile=False, symmetric=False):
  """Device-resident entries/s of one [b1, b2] block per rank (inputs in HBM before the timed region)."""
  B.set_dtype(dtype)
  kernel_fn, low = B.kernel(workload)
  H_, W_, C_ = input_dims(workload)
  ctx, lib, D = B.ctx, B.lib, B.D
  rank, world = B.rank, B.world
  names = ('nngp', 'ntk')
  out = {}
  if world > 1:
    # product path: x1 = world x b1 rows and x2 on rank 0, broadcast + row slabs + all-gather inside the step
    x1_h = np.random.default_rng(100).standard_normal(input_shape(workload, b1 * world)).astype(B.np_dt)
    x2_h = np.random.default_rng(1).standard_normal(input_shape(workload, b2)).astype(B.np_dt)
    plan = B.be.resolve(kernel_fn, x1_h.shape)
    d1 = B.be.put(x1_h, x1_h.shape, B.np_dt, 0)
    d2 = B.be.put(x2_h, x2_h.shape, B.np_dt, 0)
    held = {}

    def step():
      for v in held.values():
        B.be.free(v)
      held.clear()
      held.update(D.gram_resident(B.be, plan, d1, None if symmetric else d2, names, src=0, gather=True))
  else:
    x1_h = np.random.default_rng(100).standard_normal(input_shape(workload, b1)).astype(B.np_dt)
    x2_h = np.random.default_rng(1).standard_normal(input_shape(workload, b2)).astype(B.np_dt)
    d1, d2 = D.DeviceArray(ctx, x1_h.shape, B.np_dt), D.DeviceArray(ctx, x2_h.shape, B.np_dt)
    ctx.h2d(d1.ptr, x1_h)
    ctx.h2d(d2.ptr, x2_h)
    ctx.synchronize()
    ok, ot = D.DeviceArray(ctx, (b1, b2), B.np_dt), D.DeviceArray(ctx, (b1, b2), B.np_dt)
    held = {'nngp': ok, 'ntk': ot}

    def step():
      lib.gram_device(ctx, low.program, B.np_dt, d1.ptr, b1, None if symmetric else d2.ptr, b1 if symmetric else b2,
                      H_, W_, C_, flags, ok.ptr, ot.ptr, b2)

  for _ in range(warmup):
    step()
  ctx.synchronize()
  launches0 = ctx.launch_count
  if profile:
    ctx.set_profiling(True)
  ms = B.time_steps(step, steps, 0)
  out['launches'] = ctx.launch_count - launches0
  out['ms_dev'] = ms
  out['ms_dev_max'] = B.max_over_ranks([ms])[0]
  out['entries_per_step'] = (b1 * world) ** 2 if (symmetric and world > 1) else b1 * b2 * world
  out['value'] = out['entries_per_step'] * steps / (out['ms_dev_max'] * 1e-3)
  out['x1_h'], out['x2_h'], out['kernel_fn'] = x1_h, x2_h, kernel_fn
  out['result'] = {k: ctx.d2h(np.empty(v.shape, B.np_dt), v.ptr) for k, v in held.items()}
  for v in list(held.values()) + [d1, d2]:
    v.free()
  return out


def measure_strong(B, n, dtype, workload='myrtle10'):
  """Fixed-size symmetric Gram K(x, x) over all ranks through the public multi-GPU entry
  (`distributed.gram`: host array on rank 0 in, host matrices out on every rank), wall clock from barrier to
  barrier, max over ranks.  One untimed warm-up on a small problem, one timed run."""
  B.set_dtype(dtype)
  kernel_fn, _ = B.kernel(workload)
  be = B.be if B.be is not None else B.D.init(rank=0, world=1, local_rank=B.local, unique_id=B.lib.Comm.unique_id())
  x = np.random.default_rng(7).standard_normal(input_shape(workload, n)).astype(B.np_dt) if B.rank == 0 else None
  B.D.gram(kernel_fn, None if B.rank else x[:4 * B.world], None, ('nngp', 'ntk'), backend=be)
  be.barrier()
  t0 = time.perf_counter()
  res = B.D.gram(kernel_fn, x, None, ('nngp', 'ntk'), backend=be)
  be.barrier()
  dt = time.perf_counter() - t0
  dt_max = B.max_over_ranks([dt])[0]
  sched = B.D.sym_schedule(n, B.world, B.D.sym_block_rows(n, B.world))
  work = B.D.sym_work(sched, B.world)
  rec = {'workload': f'{workload}_32x32x3_nngp+ntk', 'n': n, 'x2': None, 'dtype': dtype, 'scaling': 'strong',
         'entries': n * n, 'pairs_computed': int(sum(work)), 'seconds': dt_max, 'entries_per_s': n * n / dt_max,
         'pairs_per_s': sum(work) / dt_max, 'imbalance_max_over_mean': max(work) / (sum(work) / B.world),
         'api': 'distributed.gram(kernel_fn, x, None) host in / host out, NCCL broadcast + all-gather + ntk_sym_assemble'}
  if B.rank == 0:
    i = np.random.default_rng(3).integers(0, n, 64)
    j = np.random.default_rng(4).integers(0, n, 64)
    rec['asymmetry_max'] = float(np.abs(res.ntk[i, j] - res.ntk[j, i]).max())
    direct = kernel_fn(x[i[:8]], x[j[:8]], ('nngp', 'ntk'))
    rec['vs_direct_block_max_rel'] = float(np.abs(res.ntk[np.ix_(i[:8], j[:8])] / direct.ntk - 1).max())
  return rec


def run_ours(args):
  B = Bench(args)
  nt, lib = B.nt, B.lib
  rank, world = B.rank, B.world
  depth, elems_net, elems_stage0 = WORKLOADS[args.workload]
  b1, b2 = block_of(args)
  flags = (lib.FLAG_NO_FUSION if args.no_fusion else 0) | (lib.FLAG_PER_LAYER if args.per_layer else 0)
  x64 = args.dtype == 'f64'
  sz = 8 if x64 else 4

  sampler = ClockSampler(B.local)
  if rank == 0:
    sampler.start()
  m = measure_block(B, args.workload, args.dtype, b1, b2, args.steps, args.warmup, flags, profile=True,
                    symmetric=args.symmetric)
  ctx = B.ctx
  ms_dev, ms_dev_max = m['ms_dev'], m['ms_dev_max']
  st0_ms, st0_n, st0_pairs = ctx.profile(0) if not args.no_fusion else (0.0, 0, 0)
  per_stage = []
  if not args.no_fusion and depth > 0:
    for si, el in enumerate(stage_elements(depth, args.per_layer)):
      ms_, n_, pr_ = ctx.profile(si)
      if n_ > 0:
        per_stage.append({'stage': si, 'ms_per_launch': ms_ / n_,
                          'algorithmic_GBps': pr_ * el * sz / (ms_ * 1e-3) / 1e9})
  ctx.set_profiling(False)

  # ---- end to end through the public API: pinned HOST buffers in, HOST results out ----------------------
  x1_h, x2_h, kernel_fn = lib.pinned_copy(m['x1_h']), lib.pinned_copy(m['x2_h']), m['kernel_fn']
  get = ('nngp', 'ntk')
  if world > 1:
    def e2e_call():
      return B.D.gram(kernel_fn, x1_h if rank == 0 else None,
                      None if args.symmetric else (x2_h if rank == 0 else None), get)
  else:
    g_ = math.gcd(b1, b2)
    e2e_cap = 500 if args.workload == 'fcn' else args.e2e_batch     # FCN entries are ~1e4 x cheaper: bigger blocks
    e2e_bs = max(d for d in range(1, min(e2e_cap, g_) + 1) if g_ % d == 0)
    batched = nt.batch(kernel_fn, batch_size=e2e_bs, device_count=0)

    def e2e_call():
      return batched(x1_h, None if args.symmetric else x2_h, get)
  for _ in range(max(1, min(args.warmup, 2))):   # untimed warm-up of the host path (IO buffers)
    e2e_call()
  B.barrier()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    res = e2e_call()
  ctx.synchronize()
  e2e_s = time.perf_counter() - t0
  B.barrier()
  e2e_ms_max = B.max_over_ranks([e2e_s * 1e3])[0]
  clocks = sampler.summary() if rank == 0 else None

  # the device-resident result of the last timed step must equal the public-API result
  if world > 1:
    np.testing.assert_allclose(m['result']['nngp'], res.nngp, rtol=1e-5)
    np.testing.assert_allclose(m['result']['ntk'], res.ntk, rtol=1e-5)
  else:
    np.testing.assert_allclose(m['result']['nngp'], res.nngp, rtol=1e-5)
    np.testing.assert_allclose(m['result']['ntk'], res.ntk, rtol=1e-5)

  strong = None
  if args.strong_n > 0 and not args.no_fusion and args.workload == 'myrtle10' and args.dtype == 'f32':
    strong = [measure_strong(B, args.strong_n, 'f32')]
    if not args.no_configs:
      # BASELINE configs[2]: Myrtle-7, 4096 x 4096 at 1 / 2 / 4 / 8 GPUs (stated size; x2 = None)
      strong.append(measure_strong(B, 4096, 'f32', 'myrtle7'))
    if world >= 8 and args.full_n > 0:
      # BASELINE configs[3] (Myrtle-10, 10000 x 10000, FP32 and FP64) and configs[4] (WideResNet Relu / Erf,
      # 4096 x 4096 at 8 GPUs) at their stated sizes
      strong.append(measure_strong(B, args.full_n, 'f32'))
      if not args.no_configs:
        strong.append(measure_strong(B, 4096, 'f32', 'wrn'))
        strong.append(measure_strong(B, 4096, 'f32', 'wrn_erf'))
        strong.append(measure_strong(B, args.full_n, 'f64'))
    B.set_dtype(args.dtype)

  configs = None
  if world == 1 and not args.no_configs and args.workload == 'myrtle10' and args.dtype == 'f32':
    configs = []
    for wl, dt, blk, stated in (('fcn', 'f32', (1000, 1000), 'configs[0]: 1000 x 1000, 784-d (stated size)'),
                                ('fcn', 'f64', (1000, 1000), 'configs[0] in float64 (stated size and dtype)'),
                                ('myrtle5', 'f32', (1024, 1024), 'configs[1]: 1024 x 1024 block, 1 B200 (stated size)'),
                                ('myrtle7', 'f32', (256, 256), 'configs[2]: one block of the 4096 x 4096 tiling'),
                                ('myrtle10', 'f64', (256, 256), 'configs[3] FP64 path: one block of the 10000 x 10000 tiling'),
                                ('wrn', 'f32', (256, 256), 'configs[4] Relu: one block of the 4096 x 4096 tiling'),
                                ('wrn', 'f64', (256, 256), 'configs[4] Relu in float64'),
                                ('wrn_erf', 'f32', (256, 256), 'configs[4] Erf: one block of the 4096 x 4096 tiling')):
      steps_c = 1 if (blk[0] >= 1024 and wl != 'fcn') or dt == 'f64' else 3
      mc = measure_block(B, wl, dt, blk[0], blk[1], steps_c, 1 if steps_c == 1 else 2)
      configs.append({'workload': workload_label(wl), 'dtype': dt, 'block': list(blk), 'steps': steps_c,
                      'ms_per_step': mc['ms_dev_max'] / steps_c, 'entries_per_s': mc['value'],
                      'gpu_launches': int(mc['launches']), 'baseline_config': stated,
                      'algorithmic_roofline_frac': (mc['value'] * WORKLOADS[wl][1] * (8 if dt == 'f64' else 4) / 1e9 /
                                                    peaks()[0]['hbm_gbs']) if WORKLOADS[wl][1] else None})
    B.set_dtype(args.dtype)
  if rank != 0:
    B.D.shutdown()
    return

  entries_per_step = m['entries_per_step']
  value = m['value']
  e2e_value = entries_per_step * args.steps / (e2e_ms_max * 1e-3)
  pk, pk_kind = peaks()
  roof = {'bound': 'hbm', 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'peak_kind': pk_kind + ' (burst copy)'}
  if st0_n > 0:
    alg_bytes_per_pair = elems_stage0 * sz
    achieved = st0_pairs * alg_bytes_per_pair / (st0_ms * 1e-3) / 1e9
    cap = ncu_capture(args.workload, args.dtype)
    tpp = cap.get('dram_bytes_per_pair') if cap else None
    n_l = 3 if depth in (10, 21) else 2
    roof.update({
        'kernel': ('k_stage<S=32,L=3,FROM_X,STORE> (first 3 fused Conv+Relu layers)' if depth == 21 else
                   ('k_stage_p' if (not x64 and args.workload != 'myrtle10_erf') else 'k_stage') +
                   '<S=32,L=%d,FROM_X,POOL> (fused Conv+%s x%d + AvgPool)' % (
                       (3, 'Erf' if args.workload == 'myrtle10_erf' else 'Relu', 3) if depth == 10 else (2, 'Relu', 2))),
        'achieved': achieved, 'frac': achieved / pk['hbm_gbs'],
        'algorithmic_bytes_per_launch': alg_bytes_per_pair * st0_pairs // st0_n,
        'avg_launch_ms': st0_ms / st0_n, 'launches_timed': st0_n,
        'share_of_step': st0_ms / ms_dev,
        'traffic': None if tpp is None else tpp * st0_pairs // st0_n,
        'traffic_source': None if cap is None else cap.get('source'),
        'whole_net_achieved': b1 * b2 * args.steps * elems_net * sz / (ms_dev * 1e-3) / 1e9,
        'per_stage': per_stage,
    })
    # The kernel's real bound: cross-layer fusion removed the per-layer HBM round trips the algorithmic model
    # counts, so it is issue / register-file bound.  Live: SM clocks per 32 element-layers and SM sub-partition
    # (one warp instruction wide), from the CUDA-event launch time and the sampled SM clock.
    sm_mhz = (clocks or {}).get('sm_mhz') or pk.get('sm_max_mhz') or 1965.0
    elem_layers = st0_pairs * float(E32) * n_l
    clk = (st0_ms * 1e-3) * sm_mhz * 1e6 * 148 * 4 / (elem_layers / 32.0)
    comp = {'bound': 'issue slots / register-file read bandwidth (not HBM)', 'clk_per_element_layer': clk,
            'sm_mhz_used': sm_mhz}
    if args.dtype in RF_FLOOR_CLK and depth in (5, 7, 10) and args.workload != 'myrtle10_erf':
      comp['register_file_floor_clk'] = RF_FLOOR_CLK[args.dtype]
      comp['frac_of_register_file_floor'] = RF_FLOOR_CLK[args.dtype] / clk
    if cap:
      for k in ('issue_slot_utilisation_pct', 'pipe_fma_pct', 'pipe_xu_pct', 'pipe_alu_pct', 'pipe_lsu_pct',
                'warp_instructions_per_element_layer', 'registers_per_thread', 'achieved_occupancy_pct',
                'dram_throughput_pct', 'warp_state_pct'):
        if k in cap:
          comp[k] = cap[k]
      comp['ncu_source'] = cap.get('source')
    roof['compute'] = comp
    if args.per_layer and per_stage:
      # one layer per launch: the dominant kernel is the slowest stage; its traffic is real
      dom = max(per_stage, key=lambda d_: d_['ms_per_launch'])
      roof.update({'kernel': 'k_stage<S=32,L=1,LOAD,STORE> (stage %d, one Conv+Relu layer)' % dom['stage'],
                   'achieved': dom['algorithmic_GBps'], 'frac': dom['algorithmic_GBps'] / pk['hbm_gbs'],
                   'avg_launch_ms': dom['ms_per_launch'], 'traffic': None})
    roof['whole_net_frac'] = roof['whole_net_achieved'] / pk['hbm_gbs']
  elif args.workload == 'fcn':
    # HBM: read x1, x2 once, write nngp + ntk; tensor pipe: the 2 b1 b2 d input GEMM (x3 passes for 3xTF32)
    bytes_step = (b1 + b2) * FCN_DIM * sz + 2 * b1 * b2 * sz
    achieved = bytes_step * args.steps / (ms_dev * 1e-3) / 1e9
    flops = 2.0 * b1 * b2 * FCN_DIM * (3 if not x64 else 1) * args.steps / (ms_dev * 1e-3)
    roof.update({'kernel': 'k_gram_tf32x3 (tcgen05 kind::tf32, 3-pass split) + Dense/Relu chain' if not x64 else
                           'k_gram_dmma (mma.sync f64) + Dense/Relu chain',
                 'achieved': achieved, 'frac': achieved / pk['hbm_gbs'], 'traffic': None,
                 'input_gemm_tflops_incl_split_passes': flops / 1e12,
                 'note': 'whole step (GEMM + 7 elementwise launches); launch-latency bound at 1000 x 1000'})
  else:
    achieved = b1 * b2 * args.steps * elems_net * sz / (ms_dev * 1e-3) / 1e9
    roof.update({'kernel': ('k_res (column-sparse residual kernels, all launches of the step)' if depth in (0, -2) else
                            'k_diagnet (diagonal column only)' if depth == -1
                            else 'per-op path (all kernels of the step)'), 'achieved': achieved,
                 'frac': achieved / pk['hbm_gbs'], 'traffic': None})
  line = {
      'metric': 'kernel_entries_per_sec', 'value': value, 'unit': 'entries/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_dev_max / args.steps,
      'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': (value / world / PUBLISHED[(args.workload, args.dtype)]
                      if (args.workload, args.dtype) in PUBLISHED else None),
      'dtype': args.dtype, 'data': 'synthetic',
      'config': make_config(args, world, b1, b2),
      'host_framework': 'NumPy + ctypes over libntk_b200.so (no torch, no jax); multi-GPU = distributed.gram_resident',
      'impl': 'ours',
      'e2e': {'value': e2e_value, 'unit': 'entries/s',
              'h2d_bytes_per_step': int(x1_h.nbytes + x2_h.nbytes),
              'd2h_bytes_per_step': int(2 * entries_per_step * sz) * (world if world > 1 else 1),
              'api': 'distributed.gram (pinned host arrays on rank 0 -> host matrices on every rank)' if world > 1
                     else 'nt.batch(kernel_fn) (pinned host arrays in, pinned host matrices out)'},
      'gpu_launches': int(m['launches']),
      'clocks': clocks,
      'roofline': roof,
  }
  if strong:
    line['strong'] = strong
  if configs:
    line['configs'] = configs
  if world == 1 and not args.no_cpu:
    cores = len(os.sched_getaffinity(0))
    pool = make_cpu_pool(cores)
    try:
      cpu_port_step(pool, args.workload, cores, 1)
      n, dt = cpu_port_step(pool, args.workload, cores, args.ref_cols)
    finally:
      pool.terminate()
    line['cpu_baseline'] = {
        'value': n / dt, 'unit': 'entries/s', 'cores': cores, 'kind': 'port',
        'sample': f'{cores} worker processes x (1 x {args.ref_cols}) pairs, NumPy float64 oracle, {dt:.1f} s'}
  emit(line)
  B.D.shutdown()


_JSON_OUT = None


def claim_stdout():
  """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner
  to fd 1 when NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the rest of the
  process and the JSON line goes to the saved descriptor."""
  global _JSON_OUT
  sys.stdout.flush()
  _JSON_OUT = os.fdopen(os.dup(1), 'w')
  os.dup2(2, 1)


def emit(line):
  out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
  out.write(json.dumps(line) + '\n')
  out.flush()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='myrtle10', choices=sorted(WORKLOADS))
  ap.add_argument('--dtype', default='f32', choices=['f32', 'f64'])
  ap.add_argument('--block', type=int, nargs=2, default=[256, 256],
                  help='pairs per rank and step: [b1, b2] (65536 pairs = one tile of the size the full-Gram jobs run)')
  ap.add_argument('--e2e-batch', type=int, default=32)
  ap.add_argument('--ref-cols', type=int, default=12,
                  help='x2 columns per worker in a CPU step (about 10 s of CPU work per step on every core)')
  ap.add_argument('--no-fusion', action='store_true')
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--no-configs', action='store_true', help='skip the per-BASELINE-config lines (N = 1)')
  ap.add_argument('--strong-n', type=int, default=2048,
                  help='size of the fixed symmetric Gram of the `strong` record (0 = skip)')
  ap.add_argument('--full-n', type=int, default=10000,
                  help='with >= 8 ranks also run BASELINE configs[3] (N x N symmetric Myrtle-10) at this size (0 = skip)')
  ap.add_argument('--symmetric', action='store_true',
                  help='time K(x1, x1) (x2=None): triangle + mirror; needs a square --block')
  ap.add_argument('--per-layer', action='store_true',
                  help='one Conv+Relu layer per kernel launch (one HBM round trip per layer)')
  args = ap.parse_args()
  claim_stdout()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
