#!/bin/bash
# host-entry tests, then bench lines (config 5, 2, 3) with the pipelined e2e
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "host or pipelined or int16 or permutation" > gpurun_out/pipe_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pipe_pytest.log
for c in 5 2 3; do
  timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/pipe_cfg$c.json 2> gpurun_out/pipe_cfg$c.err; echo "cfg $c exit $?"
  python - gpurun_out/pipe_cfg$c.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value',round(j['value']),'e2e',round(j['e2e']['value']),'blocking',round(j['e2e']['blocking_value']),'same',j['e2e']['matches_device_path'],'ms/step',round(j['ms_per_step'],2), j['clocks'])
except Exception as e: print('parse fail',e)
PY
  tail -3 gpurun_out/pipe_cfg$c.err
done
