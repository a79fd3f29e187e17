#!/bin/bash
# ncu evidence for the bench command: (1) per-launch device times of two steps, (2) one --set full capture of the
# dominant kernel (one encoder layer's four GEMM launches).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --whisper off --no-cpu-baseline --no-gpu-baseline --sustain 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 380 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 40 -c 4 -f -o gpurun_out/prof_gemm_tc2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -2 gpurun_out/ncu_full.log
