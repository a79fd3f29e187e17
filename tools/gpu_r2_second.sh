#!/bin/bash
# GPU call 2 of round 2: suite with the folded log-mel kernel, the rewritten bench, ncu of the log-mel kernels.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_tests2.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r02_tests2.log
timeout 1200 python bench.py > gpurun_out/r02_bench1.log 2>gpurun_out/r02_bench1.err; echo "bench exit $?"; tail -c 600 gpurun_out/r02_bench1.err
python - <<'PY'
import json
for ln in open("gpurun_out/r02_bench1.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        w = d.pop("whisper_large", {})
        for k in ("kernels_ms_per_step",):
            print(k, json.dumps(d.pop(k, None)))
        print(json.dumps(d))
        print("WHISPER", json.dumps(w))
PY
w=whisper_full_length
timeout 300 python tools/ncu_all.py $w > gpurun_out/ncu_all_$w.plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:logmel -f -o gpurun_out/r02_logmel python tools/ncu_all.py $w > gpurun_out/ncu_logmel.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/r02_logmel.ncu-rep --page raw --csv > gpurun_out/r02_logmel.raw.csv 2>/dev/null
