#!/bin/bash
# --set full capture of the mel kernels (batch 16, config 5 shapes); the same command runs clean first
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ARGS="--steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-gpu-baseline --allow-short-warmup --profile-steps 1 --long-file-minutes 0"
python bench.py $ARGS > gpurun_out/plain_mel.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mel_power|mel_norm" -s 2 -c 2 -o gpurun_out/prof_mel -f python bench.py $ARGS > gpurun_out/ncu_mel.log 2>&1
echo "ncu mel exit $?"; ls -la gpurun_out/prof_mel.ncu-rep
