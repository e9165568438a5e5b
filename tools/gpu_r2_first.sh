#!/bin/bash
# round 2, first call: softmax inner-loop micro-benchmark + baseline bench lines of BASELINE configs 2-4 with the round-1 kernels
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 tools/microbench/softmax_mix > gpurun_out/softmax_mix.txt 2>&1; echo "softmax_mix exit $?"
nvidia-smi --query-gpu=clocks.sm --format=csv,noheader >> gpurun_out/softmax_mix.txt
timeout 300 python bench.py --model base --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_cfg2.json 2> gpurun_out/r02a_cfg2.err; echo "cfg2 exit $?"
timeout 300 python bench.py --model small --low --res 2 --batch 256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_cfg3.json 2> gpurun_out/r02a_cfg3.err; echo "cfg3 exit $?"
timeout 400 python bench.py --model medium --low --batch 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_cfg4.json 2> gpurun_out/r02a_cfg4.err; echo "cfg4 exit $?"
tail -40 gpurun_out/softmax_mix.txt
for f in gpurun_out/r02a_cfg*.json; do python - "$f" <<PY
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1],'value',round(j['value']),'e2e',round(j['e2e']['value']),'ms/step',round(j['ms_per_step'],2),{k:round(v['ms_per_step'],2) for k,v in j['kernel_profile'].items()}, j['clocks'])
except Exception as e: print('parse fail',sys.argv[1],e)
PY
done
