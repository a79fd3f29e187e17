#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 900 python -m pytest tests/test_models_gpu.py -q -p no:cacheprovider -k "golden or variants or taps" 2>&1 | tail -2
W=${WHISPER:-off}
timeout 900 python bench.py --steps 5 --warmup 3 --whisper $W --no-cpu-baseline > gpurun_out/bench.log 2>&1
python - <<'PY'
import json
for ln in open("gpurun_out/bench.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("WavLM-L clips/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
        print(" roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "share", d["roofline"]["share_of_step"])
        for k, v in d["kernels_ms_per_step"].items(): print("   %-16s %8.3f ms  x%-3d %s" % (k, v["ms"], v["launches"], v["tflops"]))
        w = d.get("whisper_large")
        if w:
            print("Whisper-L clips/s", w["value"], "ms/step", w["ms_per_step"])
            for k, v in w["kernels_ms_per_step"].items(): print("   %-16s %8.3f ms  x%-3d %s" % (k, v["ms"], v["launches"], v["tflops"]))
PY
