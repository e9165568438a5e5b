#!/bin/bash
# one ncu --set full capture of the tensor-core kernels (after the same command ran clean without ncu)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ARGS="--steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-gpu-baseline --allow-short-warmup --profile-steps 1 --long-file-minutes 0"
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|attn_tc_kernel" -s 40 -c 6 -o gpurun_out/prof_tc python bench.py $ARGS > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
