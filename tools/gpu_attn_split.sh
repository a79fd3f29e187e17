#!/bin/bash
# A/B of the attention kernel with one (SSR_ATTN_SPLIT=1) or two softmax warps per TMEM lane quadrant.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; grep -v "Warning\|warn" gpurun_out/$name.log | tail -n ${TAILN:-30}; }
TAILN=6 run k_attn2 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k attention
for s in 2 1 2 1; do
  export SSR_ATTN_SPLIT=$s
  TAILN=3 run probe_s$s python tools/attn_probe.py wavlm whisper
done
for s in 2 1; do
  export SSR_ATTN_SPLIT=$s
  TAILN=3 run bench_s$s python bench.py --steps 5 --warmup 3 --whisper on
  python - <<PY
import json
for ln in open("gpurun_out/bench_s$s.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("split $s WavLM-L clips/s", d["value"], "ms/step", d["ms_per_step"], "parity", d.get("parity"))
        print("   attention", d["kernels_ms_per_step"].get("attention"))
        w = d.get("whisper_large")
        if w:
            print("   Whisper-L clips/s", w["value"], "ms/step", w["ms_per_step"], "parity", w.get("parity"))
            print("   attention", w["kernels_ms_per_step"].get("attention"))
PY
done
