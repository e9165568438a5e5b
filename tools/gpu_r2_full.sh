#!/bin/bash
# round 2 validation: full GPU test-suite, every BASELINE config with both baselines, launch list + per-kernel metrics
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
FULL=1 CFGS="${CFGS:-5 2 3 4}" TAG=$TAG bash tools/gpu_cfgs.sh
bash tools/gpu_ncu_all.sh
