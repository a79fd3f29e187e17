"""B200-native embedding-extraction hot path (WavLM / Whisper-encoder per-layer pooled embeddings).

The directory name mirrors the upstream repository; import it as `ssr_b200` (see /ssr_b200/__init__.py).
"""
from . import augment, pipeline  # noqa: F401
from .augment import augment_and_extract, augment_audio  # noqa: F401
from .engine import SsrError, WavLMEngine, WhisperEncoderEngine, iter_batches, shard_range  # noqa: F401
from .extract import (  # noqa: F401
    extract_embeddings_from_audio_wavlm,
    extract_embeddings_from_audio_whisper,
    extract_wavlm_embeddings,
    extract_whisper_embeddings_fixed,
    get_engine,
    invalidate_engine,
    pooled_to_layer_dict,
)

__all__ = [
    "SsrError", "WavLMEngine", "WhisperEncoderEngine", "iter_batches", "shard_range",
    "extract_wavlm_embeddings", "extract_embeddings_from_audio_wavlm", "extract_whisper_embeddings_fixed",
    "extract_embeddings_from_audio_whisper", "get_engine", "invalidate_engine", "pooled_to_layer_dict", "augment_audio",
    "augment_and_extract",
]
