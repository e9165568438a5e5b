#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "attention" -s > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?"; grep -E "attention\]|passed|failed" gpurun_out/pytest_attn.log
timeout 300 python tools/attn_trace.py 2.0 > gpurun_out/attn_trace.txt 2>&1; echo "trace exit $?"
sed -n 1,30p gpurun_out/attn_trace.txt
