#!/bin/bash
# full GPU test-suite (stop at first failure) + one bench line of the default workload without the baselines
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -${TAIL:-30} gpurun_out/pytest_quick.log
bash tools/gpu_ab.sh "WAT_DUMMY=1"
