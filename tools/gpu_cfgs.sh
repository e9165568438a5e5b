#!/bin/bash
# one bench line per BASELINE config (2, 3, 4 and optionally 5), without the CPU / torch-eager baselines unless FULL=1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
EXTRA="--no-cpu-baseline --no-gpu-baseline"
[ "$FULL" = "1" ] && EXTRA=""
for c in ${CFGS:-2 3 4}; do
  steps=5; [ $c = 2 ] && steps=20
  timeout 900 python bench.py --config $c --steps $steps --warmup 3 $EXTRA > gpurun_out/${TAG:-r02}_cfg$c.json 2> gpurun_out/${TAG:-r02}_cfg$c.err; echo "cfg$c exit $?"
  python - gpurun_out/${TAG:-r02}_cfg$c.json <<PY
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1],'value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],2),{k:round(v['ms_per_step'],2) for k,v in j['kernel_profile'].items()}, j['clocks'])
    if 'gpu_baseline' in j: print('  torch_eager',round(j['gpu_baseline']['value']),'ratio',round(j['gpu_baseline']['ours_over_baseline'],2),'materialized',round(j['gpu_baseline']['materialized_qk']['value']))
    if 'cpu_baseline' in j: print('  cpu',round(j['cpu_baseline']['value'],1))
except Exception as e: print('parse fail',sys.argv[1],e)
PY
done
