#!/bin/bash
mkdir -p gpurun_out
timeout 75 python tools/step_probe.py > gpurun_out/step_probe.log 2>&1
echo "exit $?" >> gpurun_out/step_probe.log
tail -5 gpurun_out/step_probe.log
