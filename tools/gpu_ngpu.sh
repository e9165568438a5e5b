#!/bin/bash
# N-rank bench line (torchrun), N = $1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02i_bench_${N}gpu.json 2> gpurun_out/r02i_bench_${N}gpu.err; echo "bench$N exit $?"
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r02i_bench_${N}gpu.json').read().strip().splitlines()[-1])
    print('n_gpus',j['n_gpus'],'value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],1),'gather_check',j.get('gather_check'), j['clocks'])
except Exception as e: print('parse fail',e)
PY
tail -3 gpurun_out/r02i_bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | cut -c1-300
