"""B = 1 latency anatomy: wall time per call vs the sum of device-side kernel times (profile option)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import WavLMEngine, synth  # noqa: E402

model, fe = synth.build_wavlm("large")
eng = WavLMEngine.from_hf(model, fe)
clip = synth.clip_by_index(0)
for _ in range(3):
    eng.pooled([clip])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    eng.pooled([clip])
torch.cuda.synchronize()
print(f"wall per call: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
audio = torch.from_numpy(clip[None]).cuda()
n = np.array([clip.size], np.int32)
t0 = time.perf_counter()
for _ in range(50):
    eng.pooled_device(audio, n)
t_launch = (time.perf_counter() - t0) / 50 * 1e3
torch.cuda.synchronize()
print(f"host time to enqueue one forward (async): {t_launch:.3f} ms")
eng.set_option("profile", 1)
eng.pooled_device(audio, n)
prof = eng.profile_fetch()
tot = sum(v["ms"] for v in prof.values())
nl = sum(v["launches"] for v in prof.values())
print(f"sum of device kernel times: {tot:.3f} ms over {nl} launches")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]:
    print(f"   {k:18s} {v['ms']:7.3f} ms x{v['launches']}")
