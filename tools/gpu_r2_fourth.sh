#!/bin/bash
# GPU call 4: attention — packed-pair variants and the single-block kernel for short sequences.
mkdir -p gpurun_out
timeout 300 python tools/attn_probe.py > gpurun_out/r02_attn_probe2.log 2>&1; echo "probe exit $?"; cat gpurun_out/r02_attn_probe2.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k attention > gpurun_out/r02_k4.log 2>&1; echo "kernel tests exit $?"; tail -15 gpurun_out/r02_k4.log
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -x -q -p no:cacheprovider -k "wavlm or logmel" > gpurun_out/r02_m4.log 2>&1; echo "model tests exit $?"; tail -5 gpurun_out/r02_m4.log
timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-gpu-baseline --sustain 0 > gpurun_out/r02_bench3.log 2>gpurun_out/r02_bench3.err; echo "bench exit $?"
python - <<'PY'
import json
for ln in open("gpurun_out/r02_bench3.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("wavlm", d["value"], d["ms_per_step"], {k: v["ms"] for k, v in d["kernels_ms_per_step"].items()})
        w = d["whisper_large"]
        print("whisper", w["value"], w["ms_per_step"], {k: v["ms"] for k, v in w["kernels_ms_per_step"].items()}, w["full_length"])
PY
