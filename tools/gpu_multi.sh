#!/bin/bash
# Multi-GPU evidence on N GPUs of one box (gpurun --gpus N): the bench exactly as the driver launches it (both gather
# modes), the data-parallel classifier head (BASELINE configs[3]) and the 100 k-clip sweep (configs[4]).
#   bash tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N exit $?"
timeout 600 $TR --master-port 29502 bench.py --gpus $N --steps 20 --warmup 3 --gather inline --whisper off --sustain 0 > gpurun_out/r02_bench_n${N}_inline.json 2> gpurun_out/r02_bench_n${N}_inline.err; echo "bench inline N=$N exit $?"
timeout 600 $TR --master-port 29503 tools/head_dp_check.py > gpurun_out/r02_head_dp$N.json 2> gpurun_out/r02_head_dp$N.err; echo "head exit $?"
timeout 600 $TR --master-port 29504 tools/sweep.py --model wavlm --clips 100000 > gpurun_out/r02_sweep_wavlm_n$N.json 2> gpurun_out/r02_sweep_wavlm_n$N.err; echo "sweep wavlm exit $?"
WC=100000; if [ "$N" -lt 4 ]; then WC=16384; fi
timeout 900 $TR --master-port 29505 tools/sweep.py --model whisper --clips $WC > gpurun_out/r02_sweep_whisper_n$N.json 2> gpurun_out/r02_sweep_whisper_n$N.err; echo "sweep whisper exit $?"
python - <<PY
import json
def last_json(p):
    try:
        for ln in reversed(open(p).read().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
    except Exception as e:
        return {"error": str(e)}
    return {}
N = $N
b = last_json(f"gpurun_out/r02_bench_n{N}.json")
print("bench", b.get("value"), b.get("ms_per_step"), "e2e", b.get("e2e", {}).get("value"), "sustained", b.get("sustained_value"), "gather", b.get("gather"))
w = b.get("whisper_large", {})
print("whisper", w.get("value"), w.get("ms_per_step"), "gather_ok", w.get("gather_ok"), "sustained", w.get("sustained_value"))
bi = last_json(f"gpurun_out/r02_bench_n{N}_inline.json")
print("bench inline", bi.get("value"), bi.get("ms_per_step"), bi.get("gather"))
print("head", last_json(f"gpurun_out/r02_head_dp{N}.json"))
print("sweep wavlm", last_json(f"gpurun_out/r02_sweep_wavlm_n{N}.json"))
print("sweep whisper", last_json(f"gpurun_out/r02_sweep_whisper_n{N}.json"))
PY
tail -3 gpurun_out/r02_bench_n$N.err
