#!/bin/bash
# First-contact run on a B200 box: every group in its own process + timeout so one hang cannot hide the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run k_simt      python -m pytest tests/test_kernels_gpu.py -q -k "simt" -p no:cacheprovider
run k_rowops    python -m pytest tests/test_kernels_gpu.py -q -k "layernorm or pool_mean or attention" -p no:cacheprovider
run k_tc_plain  python -m pytest tests/test_kernels_gpu.py -q -k "test_gemm_plain and tc" -p no:cacheprovider
run k_tc_rest   python -m pytest tests/test_kernels_gpu.py -q -k "(epilogue or conv_view or fused_pool) and tc" -p no:cacheprovider
