#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k attention 2>&1 | tail -3
timeout 300 python tools/attn_probe.py wavlm whisper 2>&1 | tail -4
timeout 300 python tools/attn_trace.py wavlm > gpurun_out/attn_trace_wavlm.log 2>&1
sed -n 6,9p gpurun_out/attn_trace_wavlm.log | cut -c1-250
sed -n 24,26p gpurun_out/attn_trace_wavlm.log | cut -c1-250
