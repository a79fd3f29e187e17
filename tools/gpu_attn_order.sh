#!/bin/bash
# one short GPU slot: A/B of the attention item orders
mkdir -p gpurun_out
timeout 200 python tools/attn_order_probe.py > gpurun_out/attn_order.log 2>&1
echo "exit $?" >> gpurun_out/attn_order.log
tail -20 gpurun_out/attn_order.log
