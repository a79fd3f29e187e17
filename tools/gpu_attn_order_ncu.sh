#!/bin/bash
# DRAM traffic of the attention kernel under both item orders (the probe's first 16 attention launches: 8 per order
# at the Whisper-large shape). Run only after tools/gpu_attn_order.sh exited 0.
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
  --clock-control none -k regex:attention_tc_kernel -c 16 --csv --log-file gpurun_out/attn_order_ncu.csv \
  python tools/attn_order_probe.py > gpurun_out/attn_order_ncu.log 2>&1
echo "exit $?" >> gpurun_out/attn_order_ncu.log
tail -4 gpurun_out/attn_order_ncu.log
grep -c attention_tc gpurun_out/attn_order_ncu.csv
