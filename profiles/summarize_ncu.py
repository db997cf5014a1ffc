"""Turns `ncu --page raw --csv` exports into the small summaries committed under profiles/.

  python profiles/summarize_ncu.py raw.csv [pairs_per_launch] > summary.json
"""
import csv
import json
import sys

KEYS = {
    'gpu__time_duration.sum': 'duration',
    'launch__registers_per_thread': 'registers_per_thread',
    'launch__block_size': 'block_size',
    'launch__grid_size': 'grid_size',
    'dram__bytes_read.sum': 'dram_bytes_read',
    'dram__bytes_write.sum': 'dram_bytes_write',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'achieved_occupancy_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_slot_utilisation_pct',
    'smsp__inst_executed.sum': 'warp_instructions',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_fma_pct',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_alu_pct',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'pipe_xu_pct',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active': 'pipe_lsu_pct',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_tensor_pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum': 'shared_wavefronts',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'shared_bank_conflicts',
    'sm__cycles_elapsed.max': 'sm_cycles_elapsed',
}


def to_bytes(v, unit):
  v = float(v.replace(',', ''))
  u = unit.lower()
  return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)


def main():
  rows = list(csv.reader(open(sys.argv[1])))
  pairs = int(sys.argv[2]) if len(sys.argv) > 2 else None
  hdr, units = rows[0], rows[1]
  out = []
  for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    s = {'kernel': d['Kernel Name']}
    for k, name in KEYS.items():
      if k in d and d[k] not in ('', 'n/a'):
        if name.startswith('dram_bytes'):
          s[name] = to_bytes(d[k], u[k])
        else:
          s[name] = float(d[k].replace(',', ''))
          if name == 'duration':
            s['duration_unit'] = u[k]
    st = {k.replace('smsp__pcsamp_warps_issue_stalled_', ''): float(v.replace(',', ''))
          for k, v in d.items() if 'smsp__pcsamp_warps_issue_stalled' in k and v not in ('', 'n/a')
          and not k.endswith('_not_issued')}
    tot = sum(st.values()) or 1
    s['warp_state_pct'] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
    if pairs and 'dram_bytes_read' in s:
      s['pairs_per_launch'] = pairs
      s['dram_bytes_per_pair'] = (s['dram_bytes_read'] + s['dram_bytes_write']) / pairs
    out.append(s)
  print(json.dumps(out, indent=1))


if __name__ == '__main__':
  main()
