#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; grep -v "Warning\|warn" gpurun_out/$name.log | tail -n ${TAILN:-60}; }
run m_wavlm   python -m pytest tests/test_models_gpu.py -q -rA -k "wavlm or dropin" -p no:cacheprovider
run m_whisper python -m pytest tests/test_models_gpu.py -q -rA -k "whisper or logmel" -p no:cacheprovider
