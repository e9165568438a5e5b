#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of one bench step, after the same command ran clean without ncu
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ARGS="--steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-gpu-baseline --allow-short-warmup --profile-steps 1 --long-file-minutes 0"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu1.log 2>&1
echo "ncu list exit $?"
