#!/bin/bash
# GPU call 8: host-entry pipeline (parity + e2e), attention defaults, full suite.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_tests8.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02_tests8.log
timeout 900 python bench.py --steps 20 > gpurun_out/r02_bench6.log 2>gpurun_out/r02_bench6.err; echo "bench exit $?"
python - <<'PY'
import json
for ln in open("gpurun_out/r02_bench6.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("wavlm", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["streamed_value"], d["e2e"]["streamed_ragged_value"], "sustained", d["sustained_value"], "percall", d["percall_ms"])
        print({k: v["ms"] for k, v in d["kernels_ms_per_step"].items()})
        print(d["parity"]["max_rel_err"], d["clocks"])
        w = d["whisper_large"]
        print("whisper", w["value"], w["ms_per_step"], "e2e", w["e2e"], w["streamed_value"], {k: v["ms"] for k, v in w["kernels_ms_per_step"].items()})
PY
