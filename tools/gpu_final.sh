#!/bin/bash
# Round-end verification on one B200: full GPU suite, smoke, the bench exactly as the driver runs it (both arms), then
# the ncu evidence (launch list + dominant-kernel capture + every kernel of both workloads).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.log 2>gpurun_out/bench_final.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "ref exit $?"; tail -c 700 gpurun_out/bench_ref.log
bash tools/gpu_ncu_bench.sh
bash tools/gpu_ncu_all.sh
python - <<'PY'
import json
for ln in open("gpurun_out/bench_final.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("WavLM-L clips/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"], "sustained", d.get("sustained_value"), "parity", d.get("parity", {}).get("max_rel_err"))
        print(" roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "share", d["roofline"]["share_of_step"], "clocks", d["clocks"])
        for k, v in d["kernels_ms_per_step"].items(): print("   %-16s %8.3f ms  x%-3d %s" % (k, v["ms"], v["launches"], v["tflops"]))
        w = d.get("whisper_large")
        if w:
            print("Whisper-L clips/s", w["value"], "ms/step", w["ms_per_step"], "parity", w.get("parity", {}).get("max_rel_err"), "frac", w["roofline"]["frac"], "full", w.get("full_length"))
            for k, v in w["kernels_ms_per_step"].items(): print("   %-16s %8.3f ms  x%-3d %s" % (k, v["ms"], v["launches"], v["tflops"]))
PY
