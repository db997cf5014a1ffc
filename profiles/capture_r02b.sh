# Final round-2 evidence capture (one B200): launch list of the default bench command, ncu --set full of the cross-pair
# stage kernels (fp32 packed stage 0 + stages 1, 2; fp64 stage 0), of the TMA input GEMM (tensor pipe) and of the residual
# kernel with the cp.async input ring (identity block at S = 32), WideResNet launch list, bench lines.  Outputs: gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --strong-n 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    $B > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage -s 11 -c 3 -o gpurun_out/ncu_r02_f32 \
    $B > gpurun_out/ncu_f32_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage -s 11 -c 1 -o gpurun_out/ncu_r02_f64 \
    $B --dtype f64 > gpurun_out/ncu_f64_run.log 2>&1
ncu --set full --clock-control none -k regex:k_gram_tf32x3_tma -s 2 -c 1 -o gpurun_out/ncu_r02_gemm \
    $B --workload fcn --block 8192 8192 > gpurun_out/ncu_gemm_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_res -s 2 -c 1 -o gpurun_out/ncu_r02_res \
    $B --workload wrn --block 96 96 > gpurun_out/ncu_res_run.log 2>&1
M=gpu__time_duration.sum,dram__bytes.sum,sm__inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,launch__registers_per_thread
ncu --metrics $M --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_wrn_r02.csv \
    $B --workload wrn --block 96 96 > /dev/null 2>&1
for t in f32 f64 gemm res; do
  ncu -i gpurun_out/ncu_r02_$t.ncu-rep --page raw --csv > gpurun_out/ncu_r02_${t}_raw.csv 2>/dev/null
  [ $t = f32 -o $t = res ] && ncu -i gpurun_out/ncu_r02_$t.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_r02_${t}_source.csv.gz
  rm -f gpurun_out/ncu_r02_$t.ncu-rep   # gpurun_out/ is capped at 64 MiB
done
python bench.py > gpurun_out/bench_r02_f32_n1.json 2> gpurun_out/bench_r02_f32_n1.err
python bench.py --dtype f64 --no-configs --no-cpu --strong-n 0 --steps 5 > gpurun_out/bench_r02_f64_n1.json 2>/dev/null
python bench.py --workload fcn --block 1000 1000 --no-configs --no-cpu --strong-n 0 > gpurun_out/bench_r02_fcn_f32.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2>/dev/null
ls -la gpurun_out | tail -12
