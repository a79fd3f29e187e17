"""CPU tier: the numpy oracle against (a) the golden fixtures produced by the reference's own extract_* functions
(tools/make_golden.py) and (b) the known-answer values recorded in SURVEY.md 8(c)."""
import hashlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def rel_err(got, ref):
    return float((np.abs(got - ref).max(-1) / np.abs(ref).max(-1)).max())


# ------------------------------------------------------------------------------------------------ known answers
def test_conv_lengths():
    from oracle.wavlm_oracle import CONV_KERNEL, CONV_STRIDE, num_frames

    n, seq = 48000, []
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        n = (n - k) // s + 1
        seq.append(n)
    assert seq == [9599, 4799, 2399, 1199, 599, 299, 149]
    assert num_frames(48000) == 149 and num_frames(16000) == 49 and num_frames(399) == 0 and num_frames(400) == 1


def test_bucket_function_known_answers():
    from oracle.wavlm_oracle import rel_bucket

    kat = {1: 161, -1: 1, 79: 239, -79: 79, 80: 240, 81: 240, -81: 80, 100: 247, -100: 87, 120: 254, -120: 94,
           148: 261, -148: 101}
    for rel, want in kat.items():
        assert int(rel_bucket(np.array([rel]))[0]) == want, rel
    i = np.arange(149)
    b = rel_bucket(i[None, :] - i[:, None])
    assert b.min() == 0 and b.max() == 261 and len(np.unique(b)) == 203
    lut = rel_bucket(np.arange(-148, 149)).astype(np.int16)
    assert hashlib.sha256(lut.tobytes()).hexdigest().startswith("67859a4d9be1a425")


def test_bucket_function_matches_hf_far_range():
    import torch
    from transformers.models.wavlm.modeling_wavlm import WavLMAttention

    from oracle.wavlm_oracle import rel_bucket

    att = WavLMAttention(embed_dim=64, num_heads=1)
    rel = torch.arange(-4000, 4001)
    want = att._relative_positions_bucket(rel).numpy()
    np.testing.assert_array_equal(rel_bucket(rel.numpy()), want)


def test_mel_filterbank_known_answers():
    from ssr_b200.melfilters import whisper_mel_filters

    fb = whisper_mel_filters(80)
    assert fb.shape == (201, 80) and int((fb != 0).sum()) == 391
    assert fb.max() == pytest.approx(0.025880684545274913, rel=1e-12)
    np.testing.assert_allclose(fb.sum(0)[:3], 0.02486259, rtol=1e-6)
    assert not fb[200].any()
    nz = [np.nonzero(fb[:, m])[0] for m in (0, 1, 40, 79)]
    assert [(int(a[0]), int(a[-1])) for a in nz] == [(1, 1), (1, 2), (42, 44), (186, 199)]
    assert hashlib.sha256(fb.astype(np.float32).tobytes()).hexdigest().startswith("039ce818842793b3")


def test_mel_filterbank_matches_hf():
    from transformers import WhisperFeatureExtractor

    from ssr_b200.melfilters import whisper_mel_filters

    np.testing.assert_array_equal(whisper_mel_filters(80).astype(np.float32),
                                  np.asarray(WhisperFeatureExtractor().mel_filters, dtype=np.float32))


def test_sinusoids_known_answers():
    from oracle.whisper_oracle import sinusoids

    s = sinusoids(1500, 1280)
    assert s[1, 0] == pytest.approx(0.84147096, abs=1e-7)
    assert s[1, 639] == pytest.approx(9.99999902e-05, rel=1e-5)
    assert s[1, 640] == pytest.approx(0.54030234, abs=1e-7)
    assert s[1499, 1] == pytest.approx(0.84162271, abs=2e-4)


def test_logmel_known_answers():
    from oracle.whisper_oracle import log_mel
    from ssr_b200.melfilters import whisper_mel_filters

    x = (np.random.default_rng(0).standard_normal(48000) * 0.1).astype(np.float32)
    mel = log_mel(x, whisper_mel_filters(80))
    assert mel.shape == (80, 3000)
    # padded frames sit on the clamp floor (max - 8) after (x + 4) / 4
    assert mel.min() == pytest.approx(mel.max() - 2.0, abs=1e-6)
    assert (mel[:, 400:] == mel.min()).all()
    assert -1.0 < mel.mean() < -0.9


# ------------------------------------------------------------------------------------------------ vs reference golden
@pytest.mark.parametrize("name,idx", [("tiny_stable", [0, 1, 2, 3, 4]), ("tiny_post", [0, 1, 2, 3, 4]),
                                      ("base_plus", [0, 7]), ("large", [0, 3])])
def test_wavlm_oracle_vs_reference_golden(name, idx):
    from oracle.wavlm_oracle import WavLMOracle
    from ssr_b200 import synth

    g = golden("wavlm_" + name)
    model, fe = synth.build_wavlm(name)
    assert abs(synth.state_checksum(model) - float(g["checksum"])) <= 1e-9 * float(g["checksum"])
    clips = {"base_plus": synth.noise_clips(8, 48000, seed=1234),
             "large": synth.noise_clips(2, 48000, seed=1234) + synth.mixed_clips()}.get(name, synth.mixed_clips())
    orc = WavLMOracle.from_hf(model)
    got = np.stack([orc.pooled(clips[i], fe.do_normalize) for i in idx])
    assert rel_err(got, g["pooled"][idx]) < 2e-5


def test_wavlm_oracle_matches_hf_hidden_states():
    """Hidden-state list semantics (count, which tensors) against the live HF module."""
    import torch

    from oracle.wavlm_oracle import WavLMOracle
    from ssr_b200 import synth

    for name in ("tiny_stable", "tiny_post"):
        model, fe = synth.build_wavlm(name)
        clip = synth.noise_clips(1, 20000, seed=11)[0]
        x = fe(clip, sampling_rate=16000, return_tensors="pt").input_values
        with torch.no_grad():
            ref = model(x, output_hidden_states=True).hidden_states
        hs = WavLMOracle.from_hf(model, dtype=np.float64).hidden_states(clip, fe.do_normalize)
        assert len(hs) == len(ref) == model.config.num_hidden_layers + 1
        for a, b in zip(hs, ref):
            b = b[0].numpy()
            assert a.shape == b.shape
            assert np.abs(a - b).max() <= 2e-4 * max(1.0, np.abs(b).max())


def test_whisper_oracle_vs_reference_golden():
    from oracle.whisper_oracle import WhisperEncoderOracle, log_mel
    from ssr_b200 import synth
    from ssr_b200.melfilters import whisper_mel_filters

    g = golden("whisper_tiny")
    enc, fe = synth.build_whisper_encoder("tiny")
    assert abs(synth.state_checksum(enc) - float(g["checksum"])) <= 1e-9 * float(g["checksum"])
    clips = synth.mixed_clips() + synth.noise_clips(1, 480000, seed=5)
    mf = whisper_mel_filters(80)
    orc = WhisperEncoderOracle.from_hf(enc)
    got = np.stack([orc.pooled(c, mf) for c in clips])
    assert rel_err(got, g["pooled"]) < 5e-5
    gm = golden("logmel")
    mel = np.stack([log_mel(c, mf) for c in clips])
    assert np.abs(mel[:, :, ::7] - gm["mel_sub"]).max() < 2e-5


def test_whisper_decoder_oracle_vs_reference_golden():
    """decoder_layer_* of the reference (full WhisperModel through REF extract_whisper_embeddings_fixed)."""
    from oracle.whisper_oracle import WhisperDecoderTokenOracle, WhisperEncoderOracle, log_mel
    from ssr_b200 import synth
    from ssr_b200.melfilters import whisper_mel_filters

    g = golden("whisper_tiny_full")
    model, fe = synth.build_whisper_model("tiny_full")
    assert abs(synth.state_checksum(model) - float(g["checksum"])) <= 1e-9 * float(g["checksum"])
    eo = WhisperEncoderOracle.from_hf(model.encoder)
    do = WhisperDecoderTokenOracle.from_hf(model.decoder)
    clips = synth.mixed_clips()[:4]
    mf = whisper_mel_filters(80)
    for i in (0, 3):
        hs = eo.hidden_states(log_mel(clips[i], mf))
        enc = np.stack([h.mean(0) for h in hs])
        dec = np.stack(do.hidden_states(hs[-1]))
        assert rel_err(enc[None], g["encoder"][i][None]) < 5e-5
        assert rel_err(dec[None], g["decoder"][i][None]) < 5e-5
    assert g["decoder"].shape == (4, 3, 256)


def test_whisper_oracle_vs_reference_golden_at_large_width():
    """The oracle against the reference's own output at Whisper-large's width (d = 1280, 20 heads, F = 5120; 2 + 2
    layers): encoder_layer_* and decoder_layer_* of one clip of tests/golden/whisper_wide_full.npz."""
    from oracle.whisper_oracle import WhisperDecoderTokenOracle, WhisperEncoderOracle, log_mel
    from ssr_b200 import synth
    from ssr_b200.melfilters import whisper_mel_filters

    g = golden("whisper_wide_full")
    model, fe = synth.build_whisper_model("wide_full")
    assert abs(synth.state_checksum(model) - float(g["checksum"])) <= 1e-9 * float(g["checksum"])
    assert g["encoder"].shape == (4, 3, 1280) and g["decoder"].shape == (4, 3, 1280)
    eo = WhisperEncoderOracle.from_hf(model.encoder)
    do = WhisperDecoderTokenOracle.from_hf(model.decoder)
    clip = synth.mixed_clips()[2]  # the tonal clip with a silent gap
    hs = eo.hidden_states(log_mel(clip, whisper_mel_filters(80)))
    enc = np.stack([h.mean(0) for h in hs])
    dec = np.stack(do.hidden_states(hs[-1]))
    assert rel_err(enc[None], g["encoder"][2][None]) < 5e-5
    assert rel_err(dec[None], g["decoder"][2][None]) < 5e-5


def test_golden_fixture_shapes():
    assert golden("wavlm_base_plus")["pooled"].shape == (8, 13, 768)
    assert golden("wavlm_large")["pooled"].shape == (7, 25, 1024)
    assert golden("whisper_large")["pooled"].shape == (3, 33, 1280)  # third clip: a full 30 s window
    assert golden("whisper_tiny")["pooled"].shape == (6, 3, 256)
