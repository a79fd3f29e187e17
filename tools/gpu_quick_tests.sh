#!/bin/bash
# short GPU slot: the attention / pooling kernel tests on the current build
mkdir -p gpurun_out
timeout 30 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention or pool" > gpurun_out/quick_tests4.log 2>&1
echo "exit $?" >> gpurun_out/quick_tests4.log
tail -4 gpurun_out/quick_tests4.log
