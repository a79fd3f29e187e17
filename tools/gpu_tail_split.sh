#!/bin/bash
mkdir -p gpurun_out
SSR_GEMM_TAIL_SPLIT=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "gemm" 2>&1 | tail -3
for v in 1 0 1 0; do
  SSR_GEMM_TAIL_SPLIT=$v timeout 600 python bench.py --steps 8 --warmup 3 --whisper off --no-cpu-baseline > gpurun_out/bench_ts$v.log 2>&1
  python - <<PY
import json
for ln in open("gpurun_out/bench_ts$v.log"):
    if ln.startswith("{"):
        d = json.loads(ln); k = d["kernels_ms_per_step"]
        print("split=$v clips/s", d["value"], "ms", d["ms_per_step"], "ffn2", k["gemm_ffn2"]["ms"], "out", k["gemm_out"]["ms"], "qkv", k["gemm_qkv"]["ms"], "ffn1", k["gemm_ffn1"]["ms"], "launches", d["gpu_launches"], "parity", d["parity"]["max_rel_err"])
PY
done
