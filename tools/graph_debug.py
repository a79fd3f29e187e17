"""Repeated identical small host-entry calls (eager, capture, replay) for the three model paths."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import WavLMEngine, WhisperEncoderEngine, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
clip = synth.tonal_clip(48000)
if which in ("all", "wavlm"):
    model, fe = synth.build_wavlm("tiny_stable")
    eng = WavLMEngine.from_hf(model, fe)
    outs = [eng.pooled([clip]) for _ in range(4)]
    print("wavlm ok", all(np.array_equal(outs[0], o) for o in outs), flush=True)
if which in ("all", "enc"):
    enc, wfe = synth.build_whisper_encoder("tiny")
    weng = WhisperEncoderEngine.from_hf(enc, wfe)
    outs = [weng.pooled([clip]) for _ in range(4)]
    print("whisper enc ok", all(np.array_equal(outs[0], o) for o in outs), flush=True)
if which in ("all", "full"):
    name = sys.argv[2] if len(sys.argv) > 2 else "tiny_full"
    m, f = synth.build_whisper_model(name)
    feng = WhisperEncoderEngine.from_hf(m, f)
    outs = [feng.pooled_with_decoder([clip]) for _ in range(4)]
    print("whisper full ok", all(np.array_equal(outs[0][1], o[1]) for o in outs), flush=True)
