"""Whisper decoder start-token probe: `decoder_layer_*` outputs of
REF/whisper_embeddings_large.py:257-262, 286-297 against golden fixtures produced by the reference function itself with a
full (encoder + decoder) seeded WhisperModel, and against the numpy oracle."""
import os
import wave

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
_CACHE = {}


def full(name):
    if name not in _CACHE:
        from ssr_b200 import WhisperEncoderEngine, synth

        model, fe = synth.build_whisper_model(name)
        _CACHE[name] = (model, fe, WhisperEncoderEngine.from_hf(model, fe))
    return _CACHE[name]


def check(got, ref, what, cos_min=0.9999, rel_max=1e-2):
    g, r = got.astype(np.float64), ref.astype(np.float64)
    cos = (g * r).sum(-1) / np.maximum(np.sqrt((g * g).sum(-1) * (r * r).sum(-1)), 1e-30)
    rel = np.abs(g - r).max(-1) / np.maximum(np.abs(r).max(-1), 1e-30)
    print(f"{what}: min cos {cos.min():.6f}, max rel err {rel.max():.3e}")
    assert np.isfinite(got).all() and cos.min() >= cos_min and rel.max() <= rel_max, what


# wide_full has Whisper-large's width (d=1280, 20 heads, ffn 5120): dec_xattn_kernel<5>, dec_gemv and the token
# GEMMs at D=1280 / F=5120 — the instantiations REF/whisper_embeddings_large.py:442-455 (`large`) needs
@pytest.mark.parametrize("name", ["tiny_full", "mid_full", "wide_full"])
def test_encoder_and_decoder_vs_reference_golden(name):
    from ssr_b200 import synth

    model, fe, eng = full(name)
    g = np.load(os.path.join(GOLD, f"whisper_{name}.npz"))
    assert abs(synth.state_checksum(model) - float(g["checksum"])) <= 1e-9 * float(g["checksum"]), "seeded init drifted"
    assert eng.decoder_layers == model.config.decoder_layers
    clips = synth.mixed_clips()[:4]
    enc, dec = eng.pooled_with_decoder(clips)
    assert enc.shape == g["encoder"].shape and dec.shape == g["decoder"].shape
    check(enc, g["encoder"], f"{name} encoder pooled")
    check(dec, g["decoder"], f"{name} decoder start-token states")
    # hidden_states[0] of the decoder is the (audio independent) start-token embedding
    sd = model.decoder.state_dict()
    h0 = (sd["embed_tokens.weight"][0] + sd["embed_positions.weight"][0]).numpy()
    np.testing.assert_allclose(dec[:, 0], np.broadcast_to(h0, dec[:, 0].shape), rtol=0, atol=1e-6)
    # the encoder-only entry point still agrees, and a single clip equals its in-batch result
    np.testing.assert_allclose(eng.pooled(clips), enc, rtol=0, atol=2e-6 * np.abs(enc).max())
    e1, d1 = eng.pooled_with_decoder([clips[2]])
    np.testing.assert_allclose(d1[0], dec[2], rtol=0, atol=2e-5 * np.abs(dec).max())


def test_decoder_vs_oracle_on_engine_encoder_output():
    """Decoder kernels in isolation: feed the oracle the ENGINE's own last_hidden_state (debug tap) so that only
    the decoder arithmetic differs."""
    from oracle.whisper_oracle import WhisperDecoderTokenOracle
    from ssr_b200 import synth

    model, fe, eng = full("tiny_full")
    clips = synth.mixed_clips()[:2]
    _, dec = eng.pooled_with_decoder(clips)
    last = eng.debug_fetch("last_hidden").reshape(len(clips), 1500, -1)
    orc = WhisperDecoderTokenOracle.from_hf(model.decoder)
    ref = np.stack([np.stack(orc.hidden_states(last[i])) for i in range(len(clips))])
    check(dec, ref, "decoder vs oracle (same encoder states)", cos_min=0.99998, rel_max=6e-3)


def test_dropin_returns_decoder_layers(tmp_path):
    """extract_whisper_embeddings_fixed / extract_embeddings_from_audio_whisper with a full WhisperModel return the
    reference's six keys (REF/whisper_embeddings_large.py:454-455: last three encoder and decoder layers)."""
    import ssr_b200
    from ssr_b200 import synth

    model, fe, eng = full("mid_full")
    clip = synth.tonal_clip(48000)
    p = tmp_path / "clip.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes((np.clip(clip, -1, 1) * 32767).astype("<i2").tobytes())
    ne, nd = model.config.encoder_layers + 1, model.config.decoder_layers + 1
    ei, di = [ne - 1, ne - 2, ne - 3], [nd - 1, nd - 2, nd - 3]
    out = ssr_b200.extract_whisper_embeddings_fixed(str(p), model, fe, "cuda:0", ei, di)
    assert list(out) == [f"encoder_layer_{i}" for i in ei] + [f"decoder_layer_{i}" for i in di]
    assert all(v.dtype == np.float32 and v.shape == (768,) for v in out.values())
    names = [f"encoder_layer_{ei[0]}", f"decoder_layer_{di[0]}", "decoder_layer_99"]
    quant = ssr_b200.extract.load_audio(str(p))
    out2 = ssr_b200.extract_embeddings_from_audio_whisper(quant, model, fe, "cuda", names)
    assert list(out2) == names[:2]
    for k in out2:
        np.testing.assert_allclose(out2[k], out[k], rtol=0, atol=2e-5 * max(1.0, np.abs(out[k]).max()))


def test_decoder_large_width_batched_token_gemms():
    """B > 16 sends the token-level Linear layers through the tensor-core GEMM instead of the GEMV kernel; both must
    agree with each other at Whisper-large width (the first four clips are the golden's)."""
    from ssr_b200 import synth

    model, fe, eng = full("wide_full")
    g = np.load(os.path.join(GOLD, "whisper_wide_full.npz"))
    clips = synth.mixed_clips()[:4]
    many = clips + [synth.clip_by_index(i, 30000 + 1000 * i) for i in range(16)]
    enc, dec = eng.pooled_with_decoder(many)
    check(enc[:4], g["encoder"], "wide_full encoder pooled (B=20)")
    check(dec[:4], g["decoder"], "wide_full decoder states (B=20, GEMM path)")
    _, dec_small = eng.pooled_with_decoder(clips)
    check(dec_small, dec[:4], "GEMV vs GEMM token path", cos_min=0.99999, rel_max=8e-3)
