"""TEST / BASELINE INFRASTRUCTURE ONLY — port of the reference's per-clip extraction loop for CPU timing.

The reference (REF = /root/reference, absent on the GPU box) is pure Python glue around HuggingFace `transformers`:
for every clip it calls the feature extractor, runs the model with output_hidden_states=True (batch of ONE, fp32)
and mean-pools the selected layers (REF/WavLM_embeddings.py:289-323, :583-594;
REF/whisper_embeddings_large.py:242-254, :272-283, :523-533). This module restates that glue (it does not copy it)
on top of the same third-party library, so that bench.py can time "the reference's CPU path" on the GPU box's host
cores (`cpu_baseline.kind == "port"`, and the `--impl reference` arm). It is never imported by the product package.
"""
from __future__ import annotations

import time

import numpy as np
import torch


def wavlm_extract_one(audio: np.ndarray, model, feature_extractor, layer_indices) -> dict:
    """One clip, exactly the reference's call sequence (REF/WavLM_embeddings.py:289-325)."""
    inputs = feature_extractor(audio, sampling_rate=16000, return_tensors="pt")
    with torch.no_grad():
        out = model(inputs.input_values, output_hidden_states=True, return_dict=True)
    hs = out.hidden_states
    return {f"layer_{i}": torch.mean(hs[i], dim=1).cpu().numpy().flatten() for i in layer_indices if i < len(hs)}


def whisper_extract_one(audio: np.ndarray, encoder, feature_extractor, encoder_indices) -> dict:
    """One clip, encoder part of REF/whisper_embeddings_large.py:242-254, :272-283."""
    feats = feature_extractor(audio, sampling_rate=16000, return_tensors="pt").input_features
    with torch.no_grad():
        out = encoder(feats, output_hidden_states=True, return_dict=True)
    hs = out.hidden_states
    return {f"encoder_layer_{i}": torch.mean(hs[i], dim=1).cpu().numpy().flatten()
            for i in encoder_indices if i < len(hs)}


def time_clips(fn, clips, *args) -> float:
    """Seconds to push `clips` through `fn` one by one (the reference's hot loop)."""
    t0 = time.perf_counter()
    for c in clips:
        fn(c, *args)
    return time.perf_counter() - t0
