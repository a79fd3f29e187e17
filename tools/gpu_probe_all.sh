#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/attn_probe.py
timeout 300 python tools/gemm_probe.py
