#!/bin/bash
# GPU call 3: attention kernel variants (parity + timing at both shapes), the log-mel fixes, the stage error profile.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_k3.log 2>&1; echo "kernel tests exit $?"; tail -4 gpurun_out/r02_k3.log
timeout 600 python tools/attn_probe.py > gpurun_out/r02_attn_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/r02_attn_probe.log
timeout 1500 python -m pytest tests/test_models_gpu.py -m gpu -x -q -p no:cacheprovider -s -k "logmel or whisper or error_profile" > gpurun_out/r02_m3.log 2>&1; echo "model tests exit $?"; grep -E "stage |passed|failed|Error" gpurun_out/r02_m3.log | tail -20
timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-gpu-baseline --sustain 0 > gpurun_out/r02_bench2.log 2>gpurun_out/r02_bench2.err; echo "bench exit $?"
python - <<'PY'
import json
for ln in open("gpurun_out/r02_bench2.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("wavlm", d["value"], d["ms_per_step"], {k: v["ms"] for k, v in d["kernels_ms_per_step"].items()})
        w = d["whisper_large"]
        print("whisper", w["value"], w["ms_per_step"], {k: v["ms"] for k, v in w["kernels_ms_per_step"].items()}, w["full_length"])
PY
