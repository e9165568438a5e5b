#!/bin/bash
# A/B of environment switches on the default workload (config 5), no profiler:  tools/gpu_ab.sh "VAR=0" "VAR=1" ...
# each variant = one bench.py run; prints the per-class breakdown
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-gpu-baseline ${BENCH_ARGS} > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  echo "[$v] exit $?"; python - gpurun_out/ab_$i.json <<PY
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],1),'gemm TF',round(j['roofline']['achieved']),'attn TF',round(j['roofline']['attention_tflops']), {k:round(v['ms_per_step'],2) for k,v in j['kernel_profile'].items()}, j['clocks'], {k:round(v) for k,v in j['roofline'].get('encoder_gemm_tflops',{}).items()})
except Exception as e: print('parse fail',e)
PY
  tail -3 gpurun_out/ab_$i.err
done
