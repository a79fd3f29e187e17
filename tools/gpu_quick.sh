#!/bin/bash
# kernel tests + model tests + short bench (WavLM only unless WHISPER=1)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; grep -v "Warning\|warn" gpurun_out/$name.log | tail -n ${TAILN:-30}; }
TAILN=8 run k_all python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider
TAILN=12 run m_all python -m pytest tests/test_models_gpu.py -q -p no:cacheprovider
if [ "$WHISPER" = "1" ]; then W=on; else W=off; fi
TAILN=3 run bench python bench.py --steps 5 --warmup 3 --whisper $W
python - <<'PY'
import json
for ln in open("gpurun_out/bench.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("WavLM-L clips/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity"))
        print(" roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "share", d["roofline"]["share_of_step"])
        for k, v in d["kernels_ms_per_step"].items(): print("   %-16s %8.3f ms  x%-3d %s" % (k, v["ms"], v["launches"], v["tflops"]))
        w = d.get("whisper_large")
        if w:
            print("Whisper-L clips/s", w["value"], "ms/step", w["ms_per_step"], "parity", w.get("parity"))
            for k, v in w["kernels_ms_per_step"].items(): print("   %-16s %8.3f ms  x%-3d %s" % (k, v["ms"], v["launches"], v["tflops"]))
PY
