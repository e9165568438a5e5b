#!/bin/bash
# compute-sanitizer memcheck over the smoke pass (tiny model, fp32 + bf16): after the same command ran clean without it
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck.log 2>&1
echo "sanitizer exit $?"; grep -E "ERROR SUMMARY|Invalid|smoke" gpurun_out/sanitize_memcheck.log | head -20
