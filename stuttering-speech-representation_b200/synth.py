"""Synthetic workloads shared by tests, bench.py and tools/make_golden.py: model configs (checkpoints are
unavailable offline, so weights are seeded random-init of the published architectures) and
seeded synthetic audio. `transformers` is imported lazily and only to *instantiate the reference-side model
objects* that the engine is built from — exactly what a user of the reference scripts already holds.
"""
from __future__ import annotations

import numpy as np

SR = 16000


def noise_clips(n_clips: int, n_samples: int = 48000, seed: int = 1234, sigma: float = 0.1) -> list[np.ndarray]:
    """np.random.default_rng(seed).standard_normal(n) * 0.1, one draw per clip."""
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal(n_samples).astype(np.float32) * np.float32(sigma)) for _ in range(n_clips)]


def tonal_clip(n_samples: int = 48000, f0: float = 440.0) -> np.ndarray:
    """440 Hz tone with harmonics, a silence gap and a quiet tail — exercises the Whisper (max - 8) clamp."""
    t = np.arange(n_samples, dtype=np.float64) / SR
    x = 0.3 * np.sin(2 * np.pi * f0 * t) + 0.05 * np.sin(2 * np.pi * 3 * f0 * t + 0.3)
    x[n_samples // 3: n_samples // 2] = 0.0
    x[-n_samples // 8:] *= 1e-3
    return x.astype(np.float32)


def clip_by_index(index: int, n_samples: int = 48000, seed: int = 1234) -> np.ndarray:
    """Counter-keyed clip: any sharding reproduces the same clip for a global index."""
    rng = np.random.default_rng([seed, int(index)])
    return rng.standard_normal(n_samples).astype(np.float32) * np.float32(0.1)


def mixed_clips() -> list[np.ndarray]:
    """Ragged, tonal and silent clips for edge-case parity."""
    a = noise_clips(3, 48000, seed=7)
    return [a[0], a[1][:40000], tonal_clip(48000), a[2][:16000], np.zeros(32000, np.float32)]


def aug_clips() -> list[np.ndarray]:
    """Clips for the augmentation parity cases: noise, a loud chirp that exceeds +-1 once amplified (so the final
    clamp bites), and a short clip of odd length."""
    a = noise_clips(2, 48000, seed=21)
    t = np.arange(40000, dtype=np.float64) / 16000.0
    chirp = (0.97 * np.sin(2 * np.pi * (200.0 * t + 300.0 * t * t))).astype(np.float32)
    return [a[0], chirp, a[1][:16001]]


def cluster_embeddings(n: int, D: int = 1024, n_classes: int = 8, seed: int = 0, spread: float = 5.0,
                       imbalance: float = 0.6):
    """BASELINE configs[3] stand-in for pooled embeddings with labels: Gaussian clusters in R^D, geometrically
    imbalanced class counts, per-feature offsets and scales (so that the StandardScaler matters)."""
    rng = np.random.default_rng(seed)
    # the geometry is drawn first, so that it depends on the seed only and not on n
    centres = rng.standard_normal((n_classes, D)) * spread / np.sqrt(D)
    offset = rng.standard_normal(D) * 3.0
    scale = np.exp(rng.standard_normal(D) * 0.5)
    p = imbalance ** np.arange(n_classes)
    p /= p.sum()
    y = rng.choice(n_classes, size=n, p=p)
    y[:n_classes] = np.arange(n_classes)
    X = (centres[y] + rng.standard_normal((n, D))) * scale + offset
    return X.astype(np.float32), y.astype(np.int32)


# ------------------------------------------------------------------------------------------------ model configs
def wavlm_config(name: str):
    from transformers import WavLMConfig

    if name == "base_plus":  # == WavLMConfig() defaults (HF configuration_wavlm.py:159-183)
        return WavLMConfig()
    if name == "large":
        return WavLMConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                           feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False)
    if name == "large_2l":  # WavLM-Large's shapes at 2 layers: every kernel of the Large step, few launches (ncu captures)
        return WavLMConfig(hidden_size=1024, num_hidden_layers=2, num_attention_heads=16, intermediate_size=4096,
                           feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False)
    if name == "tiny_stable":  # small pre-LN variant for fast tests
        return WavLMConfig(hidden_size=512, num_hidden_layers=3, num_attention_heads=8, intermediate_size=1024,
                           feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False)
    if name == "tiny_post":  # small post-LN / GroupNorm variant
        return WavLMConfig(hidden_size=768, num_hidden_layers=2, num_attention_heads=12, intermediate_size=1024)
    raise KeyError(name)


def wavlm_do_normalize(name: str) -> bool:
    return name in ("large", "large_2l", "tiny_stable")


def whisper_config(name: str):
    from transformers import WhisperConfig

    if name == "large":
        return WhisperConfig(d_model=1280, encoder_layers=32, encoder_attention_heads=20, encoder_ffn_dim=5120,
                             decoder_layers=32, decoder_attention_heads=20, decoder_ffn_dim=5120, num_mel_bins=80,
                             vocab_size=51865)
    if name == "large_2l":  # Whisper-large's shapes at 2 layers (ncu captures)
        return WhisperConfig(d_model=1280, encoder_layers=2, encoder_attention_heads=20, encoder_ffn_dim=5120,
                             decoder_layers=2, decoder_attention_heads=20, decoder_ffn_dim=5120, num_mel_bins=80,
                             vocab_size=51865)
    if name == "tiny":  # small config for fast tests (d_model must be a multiple of 256, head_dim 64)
        return WhisperConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=1024,
                             decoder_layers=1, decoder_attention_heads=4, decoder_ffn_dim=256, num_mel_bins=80,
                             vocab_size=1000)
    raise KeyError(name)


def build_wavlm(name: str, seed: int = 0):
    """(model, feature_extractor) as the reference holds them (REF/WavLM_embeddings.py:482-483), random init."""
    import torch
    from transformers import Wav2Vec2FeatureExtractor, WavLMModel

    torch.manual_seed(seed)
    model = WavLMModel(wavlm_config(name)).eval()
    fe = Wav2Vec2FeatureExtractor(do_normalize=wavlm_do_normalize(name))
    return model, fe


def build_whisper_encoder(name: str, seed: int = 0):
    """(encoder, feature_extractor): encoder-only instantiation of the Whisper architecture, random init."""
    import torch
    from transformers import WhisperFeatureExtractor
    from transformers.models.whisper.modeling_whisper import WhisperEncoder

    torch.manual_seed(seed)
    enc = WhisperEncoder(whisper_config(name)).eval()
    return enc, WhisperFeatureExtractor()


def whisper_full_config(name: str):
    """Configs WITH a decoder (for the decoder_layer_* outputs of REF/whisper_embeddings_large.py:286-297)."""
    from transformers import WhisperConfig

    if name == "tiny_full":
        return WhisperConfig(d_model=256, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=1024,
                             decoder_layers=2, decoder_attention_heads=4, decoder_ffn_dim=512, num_mel_bins=80,
                             vocab_size=51865)
    if name == "mid_full":
        return WhisperConfig(d_model=768, encoder_layers=3, encoder_attention_heads=12, encoder_ffn_dim=3072,
                             decoder_layers=4, decoder_attention_heads=12, decoder_ffn_dim=3072, num_mel_bins=80,
                             vocab_size=51865)
    if name == "wide_full":  # Whisper-large WIDTH (d=1280, 20 heads, ffn 5120) at 2 + 2 layers: reaches every
        # kernel instantiation the `large` branch of REF/whisper_embeddings_large.py:442-455 uses
        return WhisperConfig(d_model=1280, encoder_layers=2, encoder_attention_heads=20, encoder_ffn_dim=5120,
                             decoder_layers=2, decoder_attention_heads=20, decoder_ffn_dim=5120, num_mel_bins=80,
                             vocab_size=51865)
    raise KeyError(name)


def build_whisper_model(name: str, seed: int = 0):
    """(WhisperModel, feature_extractor) with encoder AND decoder, random init, as the reference holds it
    (REF/whisper_embeddings_large.py:431-438)."""
    import torch
    from transformers import WhisperFeatureExtractor, WhisperModel

    torch.manual_seed(seed)
    return WhisperModel(whisper_full_config(name)).eval(), WhisperFeatureExtractor()


def state_checksum(model) -> float:
    """Cheap fingerprint of seeded weights (guards golden fixtures against RNG drift)."""
    import torch

    with torch.no_grad():
        return float(sum(p.double().abs().sum() for p in model.state_dict().values() if p.dtype.is_floating_point))
