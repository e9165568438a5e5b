#!/bin/bash
# two GPUs of one box: the single-process two-device test, then the 2-rank bench line (gather_check inside)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "two_devices or hands_the_encoder" -s > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02g_bench_2gpu.json 2> gpurun_out/r02g_bench_2gpu.err; echo "bench2 exit $?"
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r02g_bench_2gpu.json').read().strip().splitlines()[-1])
    print('value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],1),'gather_check',j.get('gather_check'))
except Exception as e: print('parse fail',e)
PY
tail -3 gpurun_out/r02g_bench_2gpu.err
