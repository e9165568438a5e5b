#!/bin/bash
export PYTHONDONTWRITEBYTECODE=1
for c in 8 16 32 64; do
  python bench.py --steps 3 --warmup 3 --chunk $c --no-cpu-baseline > gpurun_out/bench_chunk$c.json 2>/dev/null
  python - <<PY
import json
j=json.loads(open('gpurun_out/bench_chunk$c.json').read().strip().splitlines()[-1])
print('chunk',$c,'value',round(j['value']),'ms/step',round(j['ms_per_step'],1),{k:round(v['ms_per_step'],1) for k,v in j['kernel_profile'].items()}, j['clocks']['sm_mhz'])
PY
done
