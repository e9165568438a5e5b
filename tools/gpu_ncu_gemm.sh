#!/bin/bash
# ncu --set full of one encoder layer's four GEMMs (qkv, out-proj, fc1, fc2) at 16 clips
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ARGS="--steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-gpu-baseline --allow-short-warmup --profile-steps 1 --long-file-minutes 0"
python bench.py $ARGS > gpurun_out/plain_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel" -s 42 -c 4 -o gpurun_out/prof_gemm python bench.py $ARGS > gpurun_out/ncu_g.log 2>&1
echo "ncu gemm exit $?"
