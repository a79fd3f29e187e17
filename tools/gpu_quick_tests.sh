#!/bin/bash
# short GPU slot: model tests other than Whisper's (those ran in the previous slot)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_models_gpu.py -x -q -k "not whisper" > gpurun_out/quick_tests3.log 2>&1
echo "exit $?" >> gpurun_out/quick_tests3.log
tail -6 gpurun_out/quick_tests3.log
