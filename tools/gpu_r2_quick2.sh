#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "${K:-gemm or layernorm or epilogue or tiny_bf16 or baseline_configs}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -${TAIL:-6} gpurun_out/pytest_quick.log
bash tools/gpu_ab.sh "$@"
