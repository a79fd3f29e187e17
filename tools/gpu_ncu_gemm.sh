#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_probe.py ffn1 > gpurun_out/probe_ffn1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/ffn1_prof python tools/gemm_probe.py ffn1 > gpurun_out/ncu_ffn1.log 2>&1
tail -3 gpurun_out/ncu_ffn1.log
