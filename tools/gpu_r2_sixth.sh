#!/bin/bash
# GPU call 6: full suite, attention item-order A/B, bench.
mkdir -p gpurun_out
timeout 300 python tools/attn_probe.py wavlm > gpurun_out/r02_attn_probe3.log 2>&1; echo "probe exit $?"; cat gpurun_out/r02_attn_probe3.log
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_tests6.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02_tests6.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-gpu-baseline --sustain 0 > gpurun_out/r02_bench5.log 2>gpurun_out/r02_bench5.err; echo "bench exit $?"
python - <<'PY'
import json
for ln in open("gpurun_out/r02_bench5.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("wavlm", d["value"], d["ms_per_step"], d["e2e"]["value"], {k: v["ms"] for k, v in d["kernels_ms_per_step"].items()})
        w = d["whisper_large"]
        print("whisper", w["value"], w["ms_per_step"], {k: v["ms"] for k, v in w["kernels_ms_per_step"].items()})
PY
