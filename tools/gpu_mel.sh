#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "mel or tiny_fp32 or transcribe or int16 or smoke" > gpurun_out/pytest_mel.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_mel.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
TAG=r02f CFGS="2 5" bash tools/gpu_cfgs.sh
bash tools/gpu_ncu_all.sh
