# A/B timing + accuracy of the NTK_B200_PVAR variants of the dominant packed kernel (stage_packed.cu).
#   0 default (degree-8 polynomial, bias-free instantiation) | 10 texture G | 11 texture G, 2 rows of lag
#   12 texture G, 3 CTAs / SM | 13 texture G, 2 rows of lag, 3 CTAs / SM
mkdir -p gpurun_out
for v in ${VARIANTS:-0 10 11 12 13}; do
  NTK_B200_PVAR=$v python bench.py --no-cpu --no-configs --strong-n 0 --steps 5 --warmup 3 2>/dev/null | \
    python -c "import json,sys; d=json.load(sys.stdin); print('variant $v', round(d['value']), 'entries/s  stage0 ms', round(d['roofline']['avg_launch_ms'],3), 'clk/elem-layer', round(d['roofline']['compute']['clk_per_element_layer'],2))"
done
python profiles/check_variant.py ${VARIANTS:-10 11 12 13}
