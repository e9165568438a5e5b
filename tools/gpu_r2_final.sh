#!/bin/bash
# end-of-round evidence: full GPU test-suite, one bench line per BASELINE config (with both baselines), the reference arm,
# the launch list and per-kernel counters of one tagging call, and a --set full capture of the tensor-core kernels
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
FULL=1 CFGS="5 2 3 4" TAG=$TAG bash tools/gpu_cfgs.sh
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> /dev/null; echo "reference arm exit $?"
ARGS="--steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-gpu-baseline --allow-short-warmup --profile-steps 1 --long-file-minutes 0"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
python bench.py $ARGS > gpurun_out/plain3.log 2>&1 &&
ncu --metrics $M --clock-control none -s 138 -c 420 --csv --log-file gpurun_out/all_kernels.csv python bench.py $ARGS > gpurun_out/ncu3.log 2>&1
echo "ncu all exit $?"
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|attn_tc_kernel|pool20_bf16|mel_power" -s 40 -c 8 -o gpurun_out/prof_tc python bench.py $ARGS > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"
