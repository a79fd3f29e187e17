"""The honest GPU baseline SURVEY.md 8(d) asks for: the stock HF modules on the same B200 (`.cuda()`, fp32 and bf16),
batched like our bench (WavLM-Large B = 256 x 3 s; Whisper-large encoder B = 64 x 30 s window), producing the same
per-layer pooled output. Device-timed with CUDA events. The CPU-side feature extractors are NOT in the timed region
(WavLM: the zero-mean / unit-variance normalisation is done on the GPU; Whisper: log-mel features are precomputed once),
which favours this baseline.

    python tools/hf_gpu_baseline.py [--steps 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import synth  # noqa: E402


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--skip-whisper", action="store_true")
    a = ap.parse_args()
    out = {"note": "stock HF transformers modules on one B200, batched, output = per-layer time-mean-pooled hidden "
                   "states; feature extraction on the CPU excluded", "torch": torch.__version__}
    B = 256
    clips = torch.from_numpy(np.stack([synth.clip_by_index(i) for i in range(B)])).cuda()
    model, fe = synth.build_wavlm("large")
    model = model.cuda().eval()
    for name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        m = model.to(dtype)
        x = clips
        if fe.do_normalize:
            x = (x - x.mean(1, keepdim=True)) / torch.sqrt(x.var(1, unbiased=False, keepdim=True) + 1e-7)
        x = x.to(dtype)

        def step():
            with torch.no_grad():
                hs = m(x, output_hidden_states=True, return_dict=True).hidden_states
                return torch.stack([h.float().mean(1) for h in hs], 1)

        try:
            ms = timed(step, a.steps)
            out[f"wavlm_large_{name}"] = {"batch": B, "ms_per_step": round(ms, 2), "clips_per_s": round(B / ms * 1e3, 1)}
        except Exception as e:  # noqa: BLE001 - e.g. out of memory
            out[f"wavlm_large_{name}"] = {"error": str(e)[:200]}
        print(json.dumps({f"wavlm_large_{name}": out[f"wavlm_large_{name}"]}), flush=True)
    del model, m
    torch.cuda.empty_cache()
    if not a.skip_whisper:
        WB = 64
        enc, wfe = synth.build_whisper_encoder("large")
        enc = enc.cuda().eval()
        feats = wfe([synth.clip_by_index(i) for i in range(WB)], sampling_rate=16000,
                    return_tensors="pt").input_features.cuda()
        for name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
            m = enc.to(dtype)
            f = feats.to(dtype)

            def wstep():
                with torch.no_grad():
                    hs = m(f, output_hidden_states=True, return_dict=True).hidden_states
                    return torch.stack([h.float().mean(1) for h in hs], 1)

            try:
                ms = timed(wstep, max(2, a.steps // 2), warmup=1)
                out[f"whisper_large_enc_{name}"] = {"batch": WB, "ms_per_step": round(ms, 2),
                                                    "clips_per_s": round(WB / ms * 1e3, 1)}
            except Exception as e:  # noqa: BLE001
                out[f"whisper_large_enc_{name}"] = {"error": str(e)[:200]}
            print(json.dumps({f"whisper_large_enc_{name}": out[f"whisper_large_enc_{name}"]}), flush=True)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
