"""Fused conv + LayerNorm + GELU kernel (option conv_ln_fused) against the two-kernel path on WavLM-Large, B = 256:
agreement of the pooled output and of the conv stack taps, device time per step, per-kernel times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import WavLMEngine, synth  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    name, B = ("tiny_stable", 5) if small else ("large", 256)
    model, fe = synth.build_wavlm(name)
    eng = WavLMEngine.from_hf(model, fe)
    clips = np.stack([synth.clip_by_index(i) for i in range(B)])
    audio = torch.from_numpy(clips).cuda()
    n = np.full(B, 48000, np.int32)
    res = {}
    for opt in (0, 1):
        eng.set_option("conv_ln_fused", opt)
        out = eng.pooled_device(audio, n).cpu().numpy()
        taps = {k: eng.debug_fetch(k) for k in ("conv1", "conv3", "conv6")}
        ms = timed(lambda: eng.pooled_device(audio, n))
        res[opt] = (out, taps, ms)
        print(f"conv_ln_fused={opt}: {ms:.3f} ms/step  {B / ms * 1e3:.1f} clips/s", flush=True)
        eng.set_option("profile", 1)
        eng.pooled_device(audio, n)
        prof = eng.profile_fetch()
        eng.set_option("profile", 0)
        for k in ("gemm_conv", "gemm_conv_ln", "layernorm_gelu"):
            if k in prof:
                print(f"     {k:16s} {prof[k]['ms']:7.3f} ms x{prof[k]['launches']}")
    a, b = res[0][0].astype(np.float64), res[1][0].astype(np.float64)
    cos = (a * b).sum(-1) / np.sqrt((a * a).sum(-1) * (b * b).sum(-1))
    rel = np.abs(a - b).max(-1) / np.abs(a).max(-1)
    print(f"pooled: min cos {cos.min():.7f}  max rel {rel.max():.3e}  finite {np.isfinite(b).all()}")
    for k in res[0][1]:
        x, y = res[0][1][k].astype(np.float64), res[1][1][k].astype(np.float64)
        print(f"{k}: max abs diff {np.abs(x - y).max():.4f}  (max |ref| {np.abs(x).max():.3f}, mean abs diff "
              f"{np.abs(x - y).mean():.2e})")


if __name__ == "__main__":
    main()
