"""Time the batched device augmentation (B = 256 x 3 s) per kind with CUDA events, and the augment -> WavLM-tiny
flow; prints achieved GB/s against algorithmic bytes (speed: 2 x (read + write) of the clip; others: read + write)."""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import augment  # noqa: E402
from ssr_b200.augment import AugOp  # noqa: E402


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    B, n = 256, 48000
    aug = augment.get_augmenter(0)
    x = torch.randn(B, n, device="cuda") * 0.1
    out = torch.empty(B, n + 8, device="cuda")
    rng = random.Random(0)
    plans = {
        "speed": [AugOp("speed", new_rate=int(16000 * rng.uniform(0.95, 1.05))) for _ in range(B)],
        "noise": [AugOp("noise", factor=0.003, seed=b) for b in range(B)],
        "volume": [AugOp("volume", factor=1.05) for _ in range(B)],
        "pitch": [AugOp("pitch", n_steps=(-2, -1, 1, 2)[b % 4]) for b in range(B)],
        "mixed": None,
    }
    random.seed(0)
    plans["mixed"] = [augment.draw_op("random") for _ in range(B)]
    for name, ops in plans.items():
        ms = timed(lambda: aug.run_device(x, [n] * B, ops, out=out))
        passes = sum(2 if op.kind == "speed" else 1 for op in ops) / B
        gb = B * n * 4 * 2 * passes / 1e9
        print(f"{name:7s}: {ms:7.3f} ms/batch  {B / ms * 1e3:10.0f} clips/s  {gb / ms * 1e3:7.1f} GB/s algorithmic",
              flush=True)
    # pitch parity against the committed reference fixture, for the record (same numbers the tests bound)
    from ssr_b200 import synth

    gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                "augment.npz"))
    clips = synth.aug_clips()
    for name in (str(s) for s in gold["pitch_names"]):
        _, ci, n_steps = name.split("/")
        random.seed(int(gold[name + "/seed"]))
        out = augment.augment_audio(clips[int(ci)].copy(), augmentation_type="pitch", variant="model_training_01")
        ref = gold[name + "/sub"]
        d = out[::4].astype(np.float64) - ref
        print(f"pitch parity {name:12s}: max abs {np.abs(d).max():.2e}  rms err / rms ref "
              f"{np.sqrt((d ** 2).mean()) / np.sqrt((ref.astype(np.float64) ** 2).mean()):.2e}")
    # CPU reference arithmetic for scale: torchaudio round trip on one clip (co-prime rate)
    import time

    import torchaudio

    xc = x[0].cpu()[None]
    t0 = time.time()
    r1 = torchaudio.transforms.Resample(16000, 16321)
    r2 = torchaudio.transforms.Resample(16321, 16000)
    y = r2(r1(xc))
    print(f"torchaudio CPU speed round trip, 1 clip, co-prime rate: {time.time() - t0:.2f} s ({y.shape[1]} samples)")


if __name__ == "__main__":
    main()
