#!/bin/bash
mkdir -p gpurun_out
WHICH=${WHICH:-whisper}
python tools/attn_probe.py $WHICH > gpurun_out/attn_probe2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -o gpurun_out/attn_prof_$WHICH python tools/attn_probe.py $WHICH > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
