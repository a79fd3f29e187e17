#!/bin/bash
# First GPU call of round 2: the GPU suite on the patched engine, a baseline bench on this box, ncu of every kernel.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02_tests.log
timeout 900 python bench.py > gpurun_out/r02_bench0.log 2>gpurun_out/r02_bench0.err; echo "bench exit $?"; tail -c 1500 gpurun_out/r02_bench0.log
bash tools/gpu_ncu_all.sh
