#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; grep -v "Warning\|warn" gpurun_out/$name.log | tail -n ${TAILN:-30}; }
run smoke python __graft_entry__.py smoke
run m_fix python -m pytest tests/test_models_gpu.py -q -rA -k "variants or dropin or whisper" -p no:cacheprovider
run bench python bench.py --steps 5 --warmup 3
