#!/bin/bash
# attention tests, then A/B of attention switches on config 5 (no profiler)
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "attention or tiny_bf16 or benchmarked" > gpurun_out/attn_ab_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/attn_ab_pytest.log
bash tools/gpu_ab.sh "$@"
