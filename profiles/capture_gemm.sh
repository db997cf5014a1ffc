# Tensor-pipe evidence for the input-layer Gram (north_star item (1)): FCN 8192 x 8192, d = 784, fp32 3xTF32.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_gram_tf32x3_pipe -s 2 -c 1 -o gpurun_out/ncu_r01_gemm \
    python bench.py --workload fcn --block 8192 8192 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_gemm_run.log 2>&1
ncu -i gpurun_out/ncu_r01_gemm.ncu-rep --page raw --csv > gpurun_out/ncu_r01_gemm_raw.csv 2>/dev/null
rm -f gpurun_out/ncu_r01_gemm.ncu-rep
python bench.py --workload fcn --block 8192 8192 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r01_fcn8192_f32.json 2>/dev/null
cat gpurun_out/bench_r01_fcn8192_f32.json | cut -c1-400
