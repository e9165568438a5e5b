#!/bin/bash
# A/B of one env switch over configs 3, 4, 2:  tools/gpu_cfg_ab.sh VAR
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for c in 3 4 2; do
  for v in 0 1; do
    env $1=$v timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --long-file-minutes 0 > gpurun_out/cab_${c}_$v.json 2> gpurun_out/cab_${c}_$v.err
    python - gpurun_out/cab_${c}_$v.json $c $v <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kp=j['kernel_profile']
    print('cfg',sys.argv[2],'flag',sys.argv[3],'value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],2),{k:round(kp[k]['ms_per_step'],2) for k in ('gemm_qkv','gemm_out','gemm_fc1','gemm_fc2','gemm_head','attention')}, j['clocks']['sm_mhz'])
except Exception as e: print('parse fail',e)
PY
  done
done
