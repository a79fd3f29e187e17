"""Generate tests/golden/*.npz by running the REFERENCE's own extraction functions (imported unmodified from
/root/reference) on seeded synthetic audio with seeded random-init HF models.

Runs only in the build container (needs /root/reference); the fixtures travel to the GPU box with the repo.
Only `load_audio` is monkeypatched (torchaudio cannot decode here and the clips are synthetic) and, for Whisper,
the decoder is a stub because the decoder pass is outside this path's scope.

    python tools/make_golden.py [--only wavlm_base_plus,...]
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def import_reference():
    """Import from a scratch cwd: the scripts create ./logs at import time (REF/WavLM_embeddings.py:16-25)."""
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="ssr_ref_"))
    sys.path.insert(0, REF)
    try:
        import WavLM_embeddings as ref_wavlm
        import whisper_embeddings_large as ref_whisper
    finally:
        os.chdir(cwd)
    return ref_wavlm, ref_whisper


def run_wavlm(ref_wavlm, name: str, clips, seed=0):
    import torch
    from ssr_b200 import synth

    model, fe = synth.build_wavlm(name, seed)
    n_hs = model.config.num_hidden_layers + 1
    table = {f"clip{i}": c for i, c in enumerate(clips)}
    ref_wavlm.load_audio = lambda path, target_sr=16000, max_length=None: table[path]
    pooled = np.zeros((len(clips), n_hs, model.config.hidden_size), np.float32)
    t0 = time.time()
    for i in range(len(clips)):
        emb = ref_wavlm.extract_wavlm_embeddings(f"clip{i}", model, fe, torch.device("cpu"), list(range(n_hs)))
        assert emb is not None and len(emb) == n_hs
        for j in range(n_hs):
            pooled[i, j] = emb[f"layer_{j}"]
    print(f"  {name}: {len(clips)} clips in {time.time() - t0:.1f}s")
    return pooled, synth.state_checksum(model)


def run_whisper(ref_whisper, name: str, clips, seed=0):
    import torch
    from ssr_b200 import synth

    enc, fe = synth.build_whisper_encoder(name, seed)
    n_hs = enc.config.encoder_layers + 1
    table = {f"clip{i}": c for i, c in enumerate(clips)}
    ref_whisper.load_audio = lambda path, target_sr=16000: table[path]
    stub_decoder = lambda **kw: types.SimpleNamespace(hidden_states=())  # noqa: E731 - decoder is out of scope
    model = types.SimpleNamespace(encoder=enc, decoder=stub_decoder)
    pooled = np.zeros((len(clips), n_hs, enc.config.d_model), np.float32)
    t0 = time.time()
    for i in range(len(clips)):
        emb = ref_whisper.extract_whisper_embeddings_fixed(f"clip{i}", model, fe, torch.device("cpu"),
                                                           list(range(n_hs)), [])
        assert emb is not None and len(emb) == n_hs
        for j in range(n_hs):
            pooled[i, j] = emb[f"encoder_layer_{j}"]
    print(f"  whisper {name}: {len(clips)} clips in {time.time() - t0:.1f}s")
    return pooled, synth.state_checksum(enc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    import torch
    from ssr_b200 import synth

    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)
    ref_wavlm, ref_whisper = import_reference()

    def want(k):
        return not only or k in only

    # BASELINE config 0: WavLM-Base+, 8 noise clips (seed 1234), FE do_normalize=False
    if want("wavlm_base_plus"):
        clips = synth.noise_clips(8, 48000, seed=1234)
        pooled, ck = run_wavlm(ref_wavlm, "base_plus", clips)
        np.savez_compressed(os.path.join(OUT, "wavlm_base_plus.npz"), pooled=pooled, checksum=ck)
    if want("wavlm_base_plus_mixed"):
        pooled, ck = run_wavlm(ref_wavlm, "base_plus", synth.mixed_clips())
        np.savez_compressed(os.path.join(OUT, "wavlm_base_plus_mixed.npz"), pooled=pooled, checksum=ck)
    # WavLM-Large: 2 noise clips of config 1's stream + the ragged / tonal / silent set
    if want("wavlm_large"):
        clips = synth.noise_clips(2, 48000, seed=1234) + synth.mixed_clips()
        pooled, ck = run_wavlm(ref_wavlm, "large", clips)
        np.savez_compressed(os.path.join(OUT, "wavlm_large.npz"), pooled=pooled, checksum=ck)
    for tiny in ("tiny_stable", "tiny_post"):
        if want("wavlm_" + tiny):
            pooled, ck = run_wavlm(ref_wavlm, tiny, synth.mixed_clips())
            np.savez_compressed(os.path.join(OUT, f"wavlm_{tiny}.npz"), pooled=pooled, checksum=ck)
    if want("whisper_tiny"):
        clips = synth.mixed_clips() + synth.noise_clips(1, 480000, seed=5)
        pooled, ck = run_whisper(ref_whisper, "tiny", clips)
        np.savez_compressed(os.path.join(OUT, "whisper_tiny.npz"), pooled=pooled, checksum=ck)
    if want("whisper_large"):
        clips = [synth.noise_clips(1, 48000, seed=1234)[0], synth.tonal_clip(48000),
                 synth.noise_clips(1, 480000, seed=5)[0]]  # the third one fills the whole 30 s window
        pooled, ck = run_whisper(ref_whisper, "large", clips)
        np.savez_compressed(os.path.join(OUT, "whisper_large.npz"), pooled=pooled, checksum=ck)
    # full Whisper models (encoder + decoder): every encoder_layer_* AND decoder_layer_* output of the reference
    for full in ("tiny_full", "mid_full", "wide_full"):
        if want("whisper_" + full):
            model, fe = synth.build_whisper_model(full, 0)
            ne, nd = model.config.encoder_layers + 1, model.config.decoder_layers + 1
            clips = synth.mixed_clips()[:4]
            table = {f"clip{i}": c for i, c in enumerate(clips)}
            ref_whisper.load_audio = lambda path, target_sr=16000: table[path]
            enc = np.zeros((len(clips), ne, model.config.d_model), np.float32)
            dec = np.zeros((len(clips), nd, model.config.d_model), np.float32)
            for i in range(len(clips)):
                emb = ref_whisper.extract_whisper_embeddings_fixed(f"clip{i}", model, fe, torch.device("cpu"),
                                                                   list(range(ne)), list(range(nd)))
                assert emb is not None and len(emb) == ne + nd
                for j in range(ne):
                    enc[i, j] = emb[f"encoder_layer_{j}"]
                for j in range(nd):
                    dec[i, j] = emb[f"decoder_layer_{j}"]
            np.savez_compressed(os.path.join(OUT, f"whisper_{full}.npz"), encoder=enc, decoder=dec,
                                checksum=synth.state_checksum(model))
            print(f"  whisper {full}: {len(clips)} clips (encoder + decoder)")
    # log-mel front end alone: HF WhisperFeatureExtractor (the call at REF/whisper_embeddings_large.py:242-246),
    # every 7th frame kept to bound the fixture size
    if want("logmel"):
        from transformers import WhisperFeatureExtractor

        fe = WhisperFeatureExtractor()
        clips = synth.mixed_clips() + synth.noise_clips(1, 480000, seed=5)
        mel = np.stack([fe(c, sampling_rate=16000, return_tensors="pt").input_features[0].numpy() for c in clips])
        np.savez_compressed(os.path.join(OUT, "logmel.npz"), mel_sub=mel[:, :, ::7].astype(np.float32),
                            stats=np.stack([mel.min((1, 2)), mel.max((1, 2)), mel.mean((1, 2))], 1))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
