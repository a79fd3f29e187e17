"""Latency of the reference-shaped per-clip calls (B = 1, host numpy in, dict of numpy out) through the drop-in
functions — what a user gets by only swapping the import, before batching anything."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssr_b200  # noqa: E402
from ssr_b200 import synth  # noqa: E402


def timed(fn, n=30):
    fn()
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    clip = synth.clip_by_index(0)
    model, fe = synth.build_wavlm("large")
    ms = timed(lambda: ssr_b200.extract_embeddings_from_audio_wavlm(clip, model, fe, "cuda", [24, 23, 22, 12]))
    print(f"WavLM-Large  per-clip drop-in call: {ms:7.2f} ms  ({1e3 / ms:7.1f} clips/s)", flush=True)
    del model
    from transformers import WhisperFeatureExtractor, WhisperModel

    torch.manual_seed(0)
    wm = WhisperModel(synth.whisper_config("large")).eval()
    wfe = WhisperFeatureExtractor()
    names = ["encoder_layer_32", "encoder_layer_31", "decoder_layer_32", "decoder_layer_31"]
    ms = timed(lambda: ssr_b200.extract_embeddings_from_audio_whisper(clip, wm, wfe, "cuda", names), n=10)
    print(f"Whisper-large per-clip drop-in call (encoder + decoder probe): {ms:7.2f} ms  ({1e3 / ms:7.1f} clips/s)")
    enc_names = names[:2]
    ms = timed(lambda: ssr_b200.extract_embeddings_from_audio_whisper(clip, wm, wfe, "cuda", enc_names), n=10)
    print(f"Whisper-large per-clip drop-in call (encoder layers only):      {ms:7.2f} ms  ({1e3 / ms:7.1f} clips/s)")


if __name__ == "__main__":
    main()
