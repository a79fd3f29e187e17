"""Drop-in replacements for the reference's four per-clip extraction functions (same names, argument order,
return type and error behaviour), backed by the CUDA engine.

    extract_wavlm_embeddings               REF/WavLM_embeddings.py:267-341
    extract_embeddings_from_audio_wavlm    REF/model_training_1.py:235-266  (== REF/model_training_01.py:214-245)
    extract_whisper_embeddings_fixed       REF/whisper_embeddings_large.py:234-299  (encoder outputs)
    extract_embeddings_from_audio_whisper  REF/model_training_1.py:268-316           (encoder outputs)

Contract kept from the reference: returns `{layer_name: float32 ndarray of shape (D,)}` in the iteration order of
the index list; an out-of-range index is skipped with a warning; any failure is logged and `None` is returned —
nothing is raised to the caller (REF/WavLM_embeddings.py:329-341).  The `model` / `feature_extractor` arguments are
the HF objects the reference scripts already hold; an engine is built from them once and cached per model object.
The `device` argument is accepted for signature compatibility: the engine always runs on CUDA.

`decoder_layer_*` entries (the reference's single start-token decoder pass) are produced when the
model object handed in has a decoder (WhisperModel / WhisperForConditionalGeneration); with a bare WhisperEncoder they
are skipped with a warning.

Batched entry points (`*_batch`) are the efficient way in: the reference's per-clip loop costs one H2D / D2H
round trip per clip.
"""
from __future__ import annotations

import logging
import wave
import weakref
from typing import Sequence

import numpy as np

from .engine import WavLMEngine, WhisperEncoderEngine

logger = logging.getLogger("ssr_b200")

_ENGINES: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _pick_params(model) -> list:
    try:
        params = list(model.parameters())
    except Exception:  # noqa: BLE001
        return []
    return params[:: max(1, len(params) // 16)]


def _weights_fingerprint(picks) -> tuple:
    """Cheap change detector for cached engines: (data pointer, version counter, dtype) of a handful of parameters
    (picked once, when the engine is built). `load_state_dict`, an optimizer step, `.half()` / `.to(dtype)` all move
    one of these."""
    return tuple((p.data_ptr(), getattr(p, "_version", 0), str(p.dtype)) for p in picks)


def get_engine(model, feature_extractor=None, device=0):
    """Engine for an HF model object, built on first use from model.state_dict() and cached per (model object,
    feature-extractor kind, CUDA device). The weights are copied at build time; if the model's parameters change
    afterwards (load_state_dict, fine-tuning, dtype cast) the fingerprint no longer matches and the engine is rebuilt.
    `invalidate_engine(model)` drops it explicitly."""
    dev = _device_index(device)
    key = (type(feature_extractor).__name__, bool(getattr(feature_extractor, "do_normalize", False)), dev)
    per_model = _ENGINES.setdefault(model, {})
    hit = per_model.get(key)
    if hit is not None:
        if _weights_fingerprint(hit[1]) == hit[2]:
            return hit[0]
        logger.warning("model parameters changed since the engine was built: rebuilding the engine")
        hit[0].close()
    name = type(model).__name__.lower()
    if "wavlm" in name:
        eng = WavLMEngine.from_hf(model, feature_extractor, dev)
    elif "whisper" in name:
        eng = WhisperEncoderEngine.from_hf(model, feature_extractor, dev)
    else:
        raise TypeError(f"unsupported model type {type(model).__name__}")
    picks = _pick_params(model)
    per_model[key] = (eng, picks, _weights_fingerprint(picks))
    return eng


def invalidate_engine(model=None) -> None:
    """Drop the cached engine(s) of `model` (all models when None): the next call rebuilds from the current weights."""
    targets = [model] if model is not None else list(_ENGINES.keys())
    for m in targets:
        for hit in _ENGINES.pop(m, {}).values():
            hit[0].close()


def _device_index(device) -> int:
    if isinstance(device, int):
        return device
    if not isinstance(device, str):
        idx = getattr(device, "index", None)  # torch.device
        if isinstance(idx, int):
            return idx
    s = str(device)
    return int(s.split(":")[1]) if ":" in s else 0


def _read_wav(file_path):
    """[channels, n] float32 + sample rate from a RIFF/WAVE file without torchaudio: integer PCM through the standard
    library, IEEE-float (format tag 3, which `wave` refuses) through a minimal chunk walk."""
    try:
        with wave.open(str(file_path), "rb") as w:
            sr, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(n)
        if width == 2:
            x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif width == 4:
            x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        elif width == 3:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
        elif width == 1:
            x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        else:
            raise ValueError(f"unsupported sample width {width}")
        return x.reshape(-1, nch).T, sr
    except wave.Error:
        pass
    import struct

    with open(str(file_path), "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file (decode it with torchaudio / convert it to WAV)")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: the real tag leads the sub-format GUID
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise ValueError("WAV file without fmt / data chunk")
    tag, nch, sr, _, _, bits = fmt
    if tag == 3 and bits in (32, 64):
        x = np.frombuffer(payload, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
        return x.reshape(-1, nch).T, sr
    raise ValueError(f"unsupported WAV format tag {tag} / {bits} bits")


def load_audio(file_path, target_sr=16000, max_length=None):
    """The reference's loader (REF/WavLM_embeddings.py:87-125; REF/whisper_embeddings_large.py has the same one without
    max_length): torchaudio.load -> mono mix-down -> torchaudio Resample to target_sr -> optional trim ->
    `.squeeze().numpy()`; any failure is logged and None returned. Where torchaudio cannot decode (this image: its
    TorchCodec backend is missing) WAV files are read by `_read_wav`; resampling still goes through torchaudio's
    Resample (pure torch), exactly as in the reference. A file that can be neither decoded nor resampled is skipped
    with an error line that says why — rows never disappear silently."""
    try:
        import torch

        try:
            import torchaudio
        except Exception:  # noqa: BLE001
            torchaudio = None
        waveform = None
        if torchaudio is not None:
            try:
                waveform, sample_rate = torchaudio.load(file_path)
            except Exception:  # noqa: BLE001 - no decoder backend: fall through to the WAV reader
                waveform = None
        if waveform is None:
            arr, sample_rate = _read_wav(file_path)
            waveform = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
        if waveform.shape[0] > 1:
            waveform = torch.mean(waveform, dim=0, keepdim=True)
        if sample_rate != target_sr:
            if torchaudio is None:
                raise ValueError(f"sample rate {sample_rate} != {target_sr} and torchaudio (Resample) is unavailable")
            waveform = torchaudio.transforms.Resample(sample_rate, target_sr)(waveform)
        if max_length is not None:
            max_samples = int(max_length * target_sr)
            if waveform.shape[1] > max_samples:
                logger.info(f"Trimming audio from {waveform.shape[1] / target_sr:.2f}s to {max_length:.2f}s")
                waveform = waveform[:, :max_samples]
        logger.debug(f"Audio shape: {waveform.shape}, duration: {waveform.shape[1] / target_sr:.2f}s")
        return np.ascontiguousarray(waveform.squeeze().numpy(), dtype=np.float32)
    except Exception as e:  # noqa: BLE001 - reference behaviour: log and return None
        logger.error(f"Error loading {file_path}: {e}")
        return None


def pooled_to_layer_dict(pooled_clip: np.ndarray, indices: Sequence[int], prefix: str) -> dict:
    """pooled_clip: [L+1, D]. Mirrors the reference's selection loop (REF/WavLM_embeddings.py:315-325)."""
    out = {}
    n = pooled_clip.shape[0]
    for idx in indices:
        if idx < n:
            out[f"{prefix}{idx}"] = np.ascontiguousarray(pooled_clip[idx], dtype=np.float32).flatten()
        else:
            logger.warning(f"Layer {idx} is out of range (max: {n - 1})")
    return out


# ---------------------------------------------------------------------------------------------- WavLM
def extract_embeddings_from_audio_wavlm(audio_array, model, feature_extractor, device, layer_indices):
    try:
        eng = get_engine(model, feature_extractor, device)
        pooled = eng.pooled([audio_array])
        return pooled_to_layer_dict(pooled[0], layer_indices, "layer_")
    except Exception as e:  # noqa: BLE001
        logger.error(f"Error extracting WavLM embeddings: {e}")
        return None


def extract_wavlm_embeddings(audio_file, model, feature_extractor, device, layer_indices, max_length=None,
                             sample_rate=16000):
    audio_array = load_audio(audio_file, target_sr=sample_rate, max_length=max_length)
    if audio_array is None:
        return None
    if audio_array.shape[0] > 500000:
        logger.warning(f"Very long input ({audio_array.shape[0]} samples, ~{audio_array.shape[0] / sample_rate:.2f}s)."
                       " This may cause memory issues.")
    return extract_embeddings_from_audio_wavlm(audio_array, model, feature_extractor, device, layer_indices)


def extract_wavlm_embeddings_batch(audio_arrays, model, feature_extractor, device, layer_indices):
    """Batched form: list of clips -> list of dicts (None for the whole batch on failure)."""
    try:
        eng = get_engine(model, feature_extractor, device)
        pooled = eng.pooled(list(audio_arrays))
        return [pooled_to_layer_dict(p, layer_indices, "layer_") for p in pooled]
    except Exception as e:  # noqa: BLE001
        logger.error(f"Error extracting WavLM embeddings: {e}")
        return None


# ---------------------------------------------------------------------------------------------- Whisper
def _whisper_states(eng, audio_array, want_decoder: bool):
    """(encoder pooled [Le+1, D], decoder start-token states [Ld+1, D] or None) for one clip. The decoder probe runs
    when the engine was built from a model that has a decoder (WhisperModel); a bare encoder yields None."""
    if want_decoder and getattr(eng, "decoder_layers", 0) > 0:
        enc, dec = eng.pooled_with_decoder([audio_array])
        return enc[0], dec[0]
    return eng.pooled([audio_array])[0], None


def extract_embeddings_from_audio_whisper(audio_array, model, processor, device, layer_names):
    try:
        eng = get_engine(model, processor, device)
        enc, dec = _whisper_states(eng, audio_array, any(n.startswith("decoder_layer_") for n in layer_names))
        out = {}
        for layer_name in layer_names:  # REF/model_training_1.py:300-312
            if layer_name.startswith("encoder_layer_"):
                idx = int(layer_name.split("_")[-1])
                if idx < enc.shape[0]:
                    out[layer_name] = np.ascontiguousarray(enc[idx], dtype=np.float32).flatten()
            elif layer_name.startswith("decoder_layer_"):
                idx = int(layer_name.split("_")[-1])
                if dec is None:
                    logger.warning(f"{layer_name}: engine was built from an encoder-only model; skipped")
                elif idx < dec.shape[0]:
                    out[layer_name] = np.ascontiguousarray(dec[idx], dtype=np.float32).flatten()
        return out
    except Exception as e:  # noqa: BLE001
        logger.error(f"Error extracting Whisper embeddings: {e}")
        return None


def extract_whisper_embeddings_fixed(audio_file, model, processor, device, encoder_indices, decoder_indices):
    audio_array = load_audio(audio_file)
    if audio_array is None:
        return None
    try:
        eng = get_engine(model, processor, device)
        enc, dec = _whisper_states(eng, audio_array, len(decoder_indices) > 0)
        out = pooled_to_layer_dict(enc, encoder_indices, "encoder_layer_")  # REF/whisper_embeddings_large.py:272-283
        if dec is not None:
            for idx in decoder_indices:  # REF/whisper_embeddings_large.py:286-297
                if idx < dec.shape[0]:
                    out[f"decoder_layer_{idx}"] = np.ascontiguousarray(dec[idx], dtype=np.float32).flatten()
                else:
                    logger.warning(f"Decoder layer {idx} is out of range (max: {dec.shape[0] - 1})")
        else:
            for idx in decoder_indices:
                logger.warning(f"decoder_layer_{idx}: engine was built from an encoder-only model; skipped")
        return out
    except Exception as e:  # noqa: BLE001
        logger.error(f"Error extracting Whisper embeddings: {e}")
        return None


def extract_whisper_embeddings_batch(audio_arrays, model, processor, device, encoder_indices):
    try:
        eng = get_engine(model, processor, device)
        pooled = eng.pooled(list(audio_arrays))
        return [pooled_to_layer_dict(p, encoder_indices, "encoder_layer_") for p in pooled]
    except Exception as e:  # noqa: BLE001
        logger.error(f"Error extracting Whisper embeddings: {e}")
        return None
