#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "benchmarked_shape or host or gemm or epilogue or baseline_configs" > gpurun_out/r02_pytest_fix.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r02_pytest_fix.log
bash tools/gpu_ab.sh "WAT_GEMM_TMA_STORE=0" "WAT_GEMM_TMA_STORE=1"
bash tools/gpu_ncu_full.sh
ls -la gpurun_out/*.ncu-rep
