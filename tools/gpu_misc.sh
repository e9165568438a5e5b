#!/bin/bash
# the less-travelled bench paths: --impl torch_eager, --config 1 (tiny, one clip, L2 flushed between steps), fp32 precision
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python bench.py --impl torch_eager --config 2 --steps 5 > gpurun_out/misc_eager_cfg2.json 2> gpurun_out/misc_eager.err; echo "eager exit $?"; cut -c1-400 gpurun_out/misc_eager_cfg2.json; tail -2 gpurun_out/misc_eager.err
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 > gpurun_out/misc_cfg1.json 2> gpurun_out/misc_cfg1.err; echo "cfg1 exit $?"; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/misc_cfg1.json').read().strip().splitlines()[-1])
    print('cfg1 value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],3),j['config']['l2'],'eager',round(j['gpu_baseline']['value']),'cpu',round(j['cpu_baseline']['value'],1), 'launches/step', j['gpu_launches']/j['steps'])
except Exception as e: print('parse fail',e)
PY
tail -2 gpurun_out/misc_cfg1.err
timeout 600 python bench.py --config 1 --precision fp32 --steps 10 --no-cpu-baseline > gpurun_out/misc_cfg1_fp32.json 2> gpurun_out/misc_cfg1_fp32.err; echo "cfg1 fp32 exit $?"; cut -c1-300 gpurun_out/misc_cfg1_fp32.json; tail -2 gpurun_out/misc_cfg1_fp32.err
