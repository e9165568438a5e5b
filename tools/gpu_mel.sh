#!/bin/bash
# mel tests + mel time at configs 5 and 2
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "mel or int16 or tiny_fp32 or host" > gpurun_out/mel_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/mel_pytest.log
for c in 5 2; do
  timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --long-file-minutes 0 > gpurun_out/mel_cfg$c.json 2> gpurun_out/mel_cfg$c.err; echo "cfg $c exit $?"
  python - gpurun_out/mel_cfg$c.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],2),'mel ms',round(j['kernel_profile']['mel']['ms_per_step'],3), j['clocks'])
except Exception as e: print('parse fail',e)
PY
  tail -3 gpurun_out/mel_cfg$c.err
done
