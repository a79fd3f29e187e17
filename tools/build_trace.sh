#!/bin/bash
# Side build of the library with the attention kernel's timeline instrumentation (tools/attn_trace.py).
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" | tail -1
B=stuttering-speech-representation_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
  -DSSR_ATT_TRACE -c stuttering-speech-representation_b200/csrc/attention_tc.cu -o $B/attention_tc_trace.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $B/libssr_trace.so $B/gemm.o $B/gemm_ln.o $B/rowops.o $B/attention.o $B/attention_tc_trace.o \
  $B/frontend.o $B/decoder.o $B/augment.o $B/head.o $B/engine.o -cudart static
echo built $B/libssr_trace.so
