#!/bin/bash
# Staged GPU check: each group in its own process (a trapped kernel poisons the CUDA context), each under a timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
export PYTHONDONTWRITEBYTECODE=1
run() { name=$1; shift; echo "=== $name" ; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -5 gpurun_out/$name.log; }
python -c "
import sys; sys.path.insert(0,'whisper-at_b200')
from whisper_at import _lib
print('tma overlap probe:', _lib.lib().wat_dbg_tma_overlap_probe())
" 2>&1 | tee gpurun_out/probe.log
run mel -k "mel"
run gemm_f32 -k "gemm_kernels and simt"
run gemm_tc -k "gemm_kernels and tcgen05"
run gemm_pair -k "gemm_kernels and ctapair"
run attn_f32 -k "attention_kernels and simt"
run attn_tc -k "attention_kernels and tcgen05"
run fp32_model -k "fp32_vs_reference or encoder_x_output or transcribe"
run bf16_model -k "tiny_low_and_base or tiny_bf16 or host_buffer or permutation"
run big -k "baseline_configs or fp32_large"
