#!/bin/bash
# GPU call 10: 128-key-block attention with P in tensor memory: parity and timing vs the 64-key kernel.
mkdir -p gpurun_out
timeout 120 python tools/attn_probe.py wavlm > gpurun_out/r02_attn_probe4w.log 2>&1; echo "probe wavlm exit $?"; cat gpurun_out/r02_attn_probe4w.log | tail -12
timeout 120 python tools/attn_probe.py whisper > gpurun_out/r02_attn_probe4h.log 2>&1; echo "probe whisper exit $?"; cat gpurun_out/r02_attn_probe4h.log | tail -12
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k attention > gpurun_out/r02_k10.log 2>&1; echo "kernel tests exit $?"; tail -12 gpurun_out/r02_k10.log
