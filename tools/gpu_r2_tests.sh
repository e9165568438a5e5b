#!/bin/bash
# full GPU test-suite, then the default bench line (config 5) with gpu_baseline and cpu_baseline
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_cfg5.json 2> gpurun_out/r02b_cfg5.err; echo "bench exit $?"
tail -3 gpurun_out/r02b_cfg5.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r02b_cfg5.json').read().strip().splitlines()[-1])
    print('value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],2),{k:round(v['ms_per_step'],2) for k,v in j['kernel_profile'].items()}, j['clocks'])
    print('gpu_baseline',json.dumps(j.get('gpu_baseline'))[:900])
    print('hbm',json.dumps(j['roofline']['hbm'])[:900])
    print('cpu',j.get('cpu_baseline'))
except Exception as e: print('parse fail',e)
PY
