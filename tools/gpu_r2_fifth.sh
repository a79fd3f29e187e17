#!/bin/bash
# GPU call 5: conv0 with analytic LayerNorm statistics (parity + timing); ncu source-level stall sampling of the two
# attention kernels (what do the warps wait for?).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -x -q -p no:cacheprovider -k "wavlm" > gpurun_out/r02_m5.log 2>&1; echo "model tests exit $?"; tail -3 gpurun_out/r02_m5.log
timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-gpu-baseline --sustain 0 --whisper off > gpurun_out/r02_bench4.log 2>gpurun_out/r02_bench4.err; echo "bench exit $?"
python - <<'PY'
import json
for ln in open("gpurun_out/r02_bench4.log"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("wavlm", d["value"], d["ms_per_step"], {k: v["ms"] for k, v in d["kernels_ms_per_step"].items()})
PY
timeout 300 python tools/attn_probe.py wavlm > gpurun_out/attn_ncu_plain_w.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_short -c 1 -f -o gpurun_out/r02_attn_short python tools/attn_probe.py wavlm > gpurun_out/attn_ncu_w.log 2>&1
echo "ncu short exit $?"
timeout 300 python tools/attn_probe.py whisper > gpurun_out/attn_ncu_plain_h.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -c 1 -f -o gpurun_out/r02_attn_whisper python tools/attn_probe.py whisper > gpurun_out/attn_ncu_h.log 2>&1
echo "ncu whisper exit $?"
ls -la gpurun_out/r02_attn_*.ncu-rep
