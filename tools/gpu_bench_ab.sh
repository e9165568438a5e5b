#!/bin/bash
# A/B: CTA-pair GEMM on/off, default workload, no profiler
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for pair in ${PAIRS:-1}; do
  WAT_GEMM_PAIR=$pair timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair$pair.json 2> gpurun_out/bench_pair$pair.err
  echo "pair=$pair exit $?"; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_pair$pair.json').read().strip().splitlines()[-1])
    print('value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],1),'gemm TF',round(j['roofline']['achieved']),'attn TF',round(j['roofline']['attention_tflops']), {k:round(v['ms_per_step'],1) for k,v in j['kernel_profile'].items()}, j['clocks'], {k:round(v) for k,v in j['roofline'].get('encoder_gemm_tflops',{}).items()})
except Exception as e: print('parse fail',e)
PY
  tail -3 gpurun_out/bench_pair$pair.err
done
