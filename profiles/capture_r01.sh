# Round-1 evidence capture (one B200): launch list, ncu --set full of the cross-pair stage kernels
# (fp32 packed stage 0 + stages 1,2; fp64 stage 0), bench lines.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage -s 11 -c 3 -o gpurun_out/ncu_r01_f32 \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_f32_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage -s 11 -c 1 -o gpurun_out/ncu_r01_f64 \
    python bench.py --steps 2 --warmup 1 --no-cpu --dtype f64 > gpurun_out/ncu_f64_run.log 2>&1
for t in f32 f64; do
  ncu -i gpurun_out/ncu_r01_$t.ncu-rep --page raw --csv > gpurun_out/ncu_r01_${t}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_r01_$t.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_r01_${t}_source.csv.gz
  rm -f gpurun_out/ncu_r01_$t.ncu-rep   # gpurun_out/ is capped at 64 MiB
done
python bench.py > gpurun_out/bench_r01_f32_n1.json 2> gpurun_out/bench_r01_f32_n1.err
python bench.py --dtype f64 --no-cpu > gpurun_out/bench_r01_f64_n1.json 2> gpurun_out/bench_r01_f64_n1.err
python bench.py --per-layer --no-cpu > gpurun_out/bench_r01_f32_perlayer.json 2> gpurun_out/bench_r01_f32_perlayer.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err
ls -la gpurun_out
