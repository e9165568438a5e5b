#!/bin/bash
# the whole -m gpu suite, then the default bench line
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 2400 python -m pytest tests -q -x -m gpu -p no:cacheprovider > gpurun_out/full_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/full_pytest.log
bash tools/gpu_ab.sh "$@"
