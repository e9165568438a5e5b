#!/bin/bash
# head / resolution tests, then config 3 and 5 bench lines
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "tltr or tiny or baseline_configs or transcribe or permutation or benchmarked" > gpurun_out/head_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/head_pytest.log
for c in 3 5 4; do
  timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --long-file-minutes 0 > gpurun_out/head_cfg$c.json 2> gpurun_out/head_cfg$c.err; echo "cfg $c exit $?"
  python - gpurun_out/head_cfg$c.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kp=j['kernel_profile']
    print('value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],2),'head_attn',round(kp['head_attention']['ms_per_step'],3),'gemm_head',round(kp['gemm_head']['ms_per_step'],2),'mean',round(kp['mean']['ms_per_step'],2), j['clocks'])
except Exception as e: print('parse fail',e)
PY
  tail -3 gpurun_out/head_cfg$c.err
done
