#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; grep -v "Warning\|warn" gpurun_out/$name.log | tail -n ${TAILN:-30}; }
TAILN=25 run k_gemm python -m pytest tests/test_kernels_gpu.py -q -k "gemm or layernorm" -p no:cacheprovider -x
TAILN=8 run probe python tools/gemm_probe.py
