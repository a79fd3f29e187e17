#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; grep -v "Warning\|warn" gpurun_out/$name.log | tail -n ${TAILN:-30}; }
run k_attn python -m pytest tests/test_kernels_gpu.py -q -k "attention" -p no:cacheprovider
TAILN=12 run m_all python -m pytest tests/test_models_gpu.py -q -p no:cacheprovider
TAILN=3 run bench python bench.py --steps 5 --warmup 3
