"""Device-timed step and per-kernel-group breakdown of both bench workloads on the current build (short GPU slots:
no CPU baseline, no oracle parity — bench.py is the measurement of record, this is its timed region alone)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ssr_b200 import WavLMEngine, WhisperEncoderEngine, synth

T0 = time.time()
out = {}


def timed(eng, clips, n, steps):
    audio = torch.from_numpy(clips).cuda()
    o = eng.pooled_device(audio, n)
    for _ in range(3):
        o = eng.pooled_device(audio, n, out=o)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        o = eng.pooled_device(audio, n, out=o)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps


def leg(name, eng, clips, n, steps):
    ms = timed(eng, clips, n, steps)
    prof = bench.profile_engine(eng, clips, n, 0, reps=1)
    out[name] = {"ms_per_step": round(ms, 3), "clips_per_s": round(len(clips) / ms * 1e3, 1), "steps": steps,
                 "kernels_ms_per_step": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}
    print(f"[{time.time() - T0:5.1f}s]", name, json.dumps(out[name]), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "step_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


enc, wfe = synth.build_whisper_encoder("large", seed=0)
weng = WhisperEncoderEngine.from_hf(enc, wfe, device=0)
wclips = np.stack([synth.clip_by_index(i, 48000) for i in range(64)])
leg("whisper_large_b64", weng, wclips, np.full(64, 48000, np.int32), 6)
del weng, enc
torch.cuda.empty_cache()
model, fe = synth.build_wavlm("large", seed=0)
eng = WavLMEngine.from_hf(model, fe, device=0)
clips = np.stack([synth.clip_by_index(i, 48000) for i in range(256)])
leg("wavlm_large_b256", eng, clips, np.full(256, 48000, np.int32), 20)
