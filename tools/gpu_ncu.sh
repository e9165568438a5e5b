#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ARGS="--steps 1 --warmup 1 --batch 16 --no-cpu-baseline --allow-short-warmup"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu1.log 2>&1
echo "ncu list exit $?"
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|attn_tc_kernel" -s 40 -c 6 -o gpurun_out/prof_tc python bench.py $ARGS > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
