#!/bin/bash
mkdir -p gpurun_out
python tools/attn_probe.py > gpurun_out/attn_probe.log 2>&1 && cat gpurun_out/attn_probe.log &&
python tools/attn_probe.py wavlm > gpurun_out/attn_probe2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -o gpurun_out/attn_prof python tools/attn_probe.py wavlm > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
