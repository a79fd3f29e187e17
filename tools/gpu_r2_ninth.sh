#!/bin/bash
# GPU call 9: programmatic dependent launch A/B, full suite.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_tests9.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r02_tests9.log
for pdl in 1 0 1 0; do
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline --sustain 0 --tune pdl=$pdl > gpurun_out/r02_bench_pdl$pdl.log 2>gpurun_out/r02_bench_pdl$pdl.err; echo "bench pdl=$pdl exit $?"
python - <<PY
import json
for ln in open("gpurun_out/r02_bench_pdl$pdl.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        w = d["whisper_large"]
        print("pdl=$pdl wavlm", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "percall", d.get("percall_ms"), "| whisper", w["value"], w["ms_per_step"], "e2e", w["e2e"])
PY
done
