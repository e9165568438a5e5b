#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ARGS="--steps 1 --warmup 1 --batch 8 --no-cpu-baseline --allow-short-warmup"
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 34 -c 1 -o gpurun_out/prof_attn python bench.py $ARGS > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
