#!/bin/bash
# ncu --set full over every kernel of one forward of both workloads (2-layer models of the published widths).
# The raw pages are exported to CSV on the box (small); a .ncu-rep is only kept when it is small enough to travel.
mkdir -p gpurun_out
for w in wavlm whisper whisper_full_length; do
  timeout 300 python tools/ncu_all.py $w > gpurun_out/ncu_all_$w.plain.log 2>&1 || { echo "$w plain run failed"; tail -5 gpurun_out/ncu_all_$w.plain.log; continue; }
  extra=""
  # the full-length run only adds the front end (log-mel, conv stem): the encoder kernels are the same as `whisper`
  if [ $w = whisper_full_length ]; then extra="-k regex:logmel|mel"; fi
  timeout 1500 ncu --set full --clock-control none --profile-from-start off $extra \
      -f -o gpurun_out/r02_all_$w python tools/ncu_all.py $w > gpurun_out/ncu_all_$w.log 2>&1
  echo "$w ncu exit $?"; tail -2 gpurun_out/ncu_all_$w.log
  ncu -i gpurun_out/r02_all_$w.ncu-rep --page raw --csv > gpurun_out/r02_all_$w.raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/r02_all_$w.ncu-rep 2>/dev/null || echo 0)
  echo "$w rep bytes $sz"
  if [ "$sz" -gt 16000000 ]; then rm -f gpurun_out/r02_all_$w.ncu-rep; fi
done
ls -la gpurun_out/r02_all_*
