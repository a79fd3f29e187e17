"""Time the Whisper-large encoder alone vs encoder + decoder start-token probe (B = 64, 3 s clips) with CUDA events.
Builds the full random-init WhisperModel (1.5 B parameters), so it is not part of the default bench."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import WhisperEncoderEngine, synth  # noqa: E402


def main():
    from transformers import WhisperFeatureExtractor, WhisperModel

    t0 = time.time()
    cfg = synth.whisper_config("large")
    torch.manual_seed(0)
    model = WhisperModel(cfg).eval()
    eng = WhisperEncoderEngine.from_hf(model, WhisperFeatureExtractor())
    print(f"model + engine built in {time.time() - t0:.1f}s, decoder layers {eng.decoder_layers}", flush=True)
    B = 64
    clips = np.stack([synth.clip_by_index(i) for i in range(B)])
    n = np.full(B, 48000, np.int32)
    audio = torch.from_numpy(clips).cuda()
    enc = torch.empty((B, 33, 1280), device="cuda")
    dec = torch.empty((B, eng.decoder_layers + 1, 1280), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib = eng._lib
    from ssr_b200 import _lib

    def full():
        rc = lib.ssr_whisper_full(eng._h, audio.data_ptr(), audio.stride(0), n.ctypes.data_as(_lib.c_i32p), B,
                                  enc.data_ptr(), dec.data_ptr(), st)
        assert rc == 0, eng._err()

    def enc_only():
        eng.pooled_device(audio, n, out=enc)

    for name, fn in (("encoder only", enc_only), ("encoder + decoder probe", full)):
        fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(3):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 3
        print(f"{name:26s}: {ms:8.2f} ms/step  {B / ms * 1e3:7.1f} clips/s", flush=True)
    eng.set_option("profile", 1)
    full()
    prof = eng.profile_fetch()
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]:
        print(f"   {k:22s} {v['ms']:8.3f} ms x{v['launches']}")
    assert np.isfinite(dec.cpu().numpy()).all()


if __name__ == "__main__":
    main()
