# Round-1 bench sweep: every BASELINE config family on one B200 (device-resident value + e2e).
mkdir -p gpurun_out
for w in "myrtle10 f32" "myrtle10 f64" "myrtle5 f32" "myrtle7 f32" "myrtle10_erf f32" "wrn f32" "wrn f64" "wrn_erf f32" "readme21 f32" "readme21_flatten f32"; do
  set -- $w
  python bench.py --workload $1 --dtype $2 --no-cpu --steps 5 --warmup 3 > gpurun_out/sweep_$1_$2.json 2> gpurun_out/sweep_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
  d = json.load(open(f'gpurun_out/sweep_{sys.argv[1]}_{sys.argv[2]}.json'))
  r = d['roofline']
  print(f"{sys.argv[1]:18s} {sys.argv[2]}  value {d['value']:12.1f}  e2e {d['e2e']['value']:12.1f}  frac {r.get('whole_net_frac', r.get('frac')):.3f}  launches {d['gpu_launches']}")
except Exception as e:
  print(sys.argv[1], sys.argv[2], 'FAILED', e)
PY
done
