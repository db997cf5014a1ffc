mkdir -p gpurun_out
for v in 0 2; do   # 0 = default, 2 = 3 CTAs per SM + planar q2 (NTK_B200_PVAR, stage_packed.cu)
  echo "variant $v"; NTK_B200_PVAR=$v python bench.py --no-cpu --steps 5 --warmup 2 | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['roofline']['per_stage'][0])"
done
