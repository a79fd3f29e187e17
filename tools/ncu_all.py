"""One forward of every kernel of the two BASELINE workloads, for `ncu --set full` (tools/gpu_ncu_all.sh).

The models have the published WIDTHS (WavLM-Large d=1024 / F=4096 / 16 heads at B=256 x 3 s; Whisper-large d=1280 /
F=5120 / 20 heads at B=64 x 30 s window) but only 2 transformer layers, so that a `--set full` capture of EVERY launch
of one forward stays short: the per-layer kernels see exactly the shapes of the full models. The profiled region is
bracketed with cudaProfilerStart/Stop (run ncu with --profile-from-start off).

    python tools/ncu_all.py wavlm|whisper|whisper_full_length
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ssr_b200 import WavLMEngine, WhisperEncoderEngine, synth  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "wavlm"
    if what == "wavlm":
        model, fe = synth.build_wavlm("large_2l", seed=0)
        eng = WavLMEngine.from_hf(model, fe)
        B, n = 256, 48000
    else:
        enc, fe = synth.build_whisper_encoder("large_2l", seed=0)
        eng = WhisperEncoderEngine.from_hf(enc, fe)
        B, n = 64, (480000 if what == "whisper_full_length" else 48000)
    clips = np.stack([synth.clip_by_index(i, n) for i in range(B)])
    audio = torch.from_numpy(clips).cuda()
    ns = np.full(B, n, np.int32)
    for _ in range(2):
        eng.pooled_device(audio, ns)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.pooled_device(audio, ns)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("done", what)


if __name__ == "__main__":
    main()
