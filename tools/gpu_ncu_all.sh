#!/bin/bash
# per-launch ncu metrics (not --set full: a handful of counters, few replay passes) for EVERY kernel of one tagging call:
# mel, conv stem, 32 encoder layers, pooling, TL-TR head.  Runs after the same command exited 0 without ncu.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ARGS="--steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-gpu-baseline --allow-short-warmup --profile-steps 1 --long-file-minutes 0"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
python bench.py $ARGS > gpurun_out/plain3.log 2>&1 &&
ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/all_kernels.csv python bench.py $ARGS > gpurun_out/ncu3.log 2>&1
echo "ncu all exit $?"
