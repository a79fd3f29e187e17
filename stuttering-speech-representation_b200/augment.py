"""Batched waveform augmentation on the GPU feeding the extraction engine.

Mirrors the reference's second caller of the hot path:

    augment_audio              REF/model_training_1.py:166-213   (variant "model_training_1": speed/noise/volume/none)
    augment_audio              REF/model_training_01.py:140-192  (variant "model_training_01": wider ranges, + pitch)
    augment_minority_classes   REF/model_training_1.py:318-430   (inner loop :378-393: augment -> extract, B = 1)

Split of work: *which* augmentation and *which* factor is decided on the host, consuming python's `random` stream
in exactly the reference's order (`draw_op`), so a script seeded like the reference makes the same decisions. The
arithmetic (sinc resampling round trip, noise add, gain, clamp) runs on the device through the C ABI `ssr_augment`,
batched, and its output stays on the device for the encoder — the reference instead re-extracts one clip at a time.

`augment_audio` keeps the reference's signature and "on failure log and return the input" convention. The noise
kind draws its normals with `torch.randn_like` on the CPU generator, like the reference, so that output is
bit-identical under the same seeds; the batched API defaults to the device generator (Philox, keyed by clip).

The `pitch` kind of model_training_01 (torchaudio PitchShift: STFT -> phase vocoder -> inverse STFT -> resample) is
built as well; being a phase vocoder its output is only statistically reproducible on tonal input (the reference's own
output moves at the percent level with FFT rounding), see csrc/augment.cu.
"""
from __future__ import annotations

import ctypes as C
import logging
import random
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .engine import SsrError, _as_f32_clip, row_pitch

logger = logging.getLogger("ssr_b200")

VARIANTS = {
    "model_training_1": {"kinds": ["speed", "noise", "volume", "none"], "speed": (0.95, 1.05),
                         "noise": (0.001, 0.005), "volume": (0.9, 1.1)},
    "model_training_01": {"kinds": ["speed", "noise", "pitch", "volume"], "speed": (0.9, 1.1),
                          "noise": (0.005, 0.02), "volume": (0.8, 1.2)},
}
_KIND_CODE = {"none": _lib.SSR_AUG_NONE, "speed": _lib.SSR_AUG_SPEED, "noise": _lib.SSR_AUG_NOISE,
              "volume": _lib.SSR_AUG_VOLUME, "pitch": _lib.SSR_AUG_PITCH}


@dataclass
class AugOp:
    kind: str = "none"
    new_rate: int = 0
    factor: float = 0.0
    n_steps: int = 0
    seed: int = 0


def draw_op(augmentation_type="random", sample_rate=16000, variant="model_training_1", rng=random) -> AugOp:
    """One decision, drawn like REF/model_training_1.py:179-199 (choice, then one uniform)."""
    v = VARIANTS[variant]
    kind = augmentation_type
    if kind == "random":
        kind = rng.choice(v["kinds"])
    if kind == "speed":
        speed_factor = rng.uniform(*v["speed"])
        return AugOp("speed", new_rate=int(sample_rate * speed_factor))
    if kind == "noise":
        return AugOp("noise", factor=rng.uniform(*v["noise"]))
    if kind == "volume":
        return AugOp("volume", factor=rng.uniform(*v["volume"]))
    if kind == "pitch":
        return AugOp("pitch", n_steps=rng.randint(-2, 2))
    return AugOp("none")


def out_length(op: AugOp, n: int, sample_rate=16000) -> int:
    c = _lib.AugOp(kind=_KIND_CODE[op.kind], new_rate=op.n_steps if op.kind == "pitch" else op.new_rate,
                   factor=op.factor, reserved=0, seed=op.seed)
    return int(_lib.load().ssr_augment_out_length(C.byref(c), int(n), int(sample_rate)))


class Augmenter:
    """Device-side batch augmenter. Owns a grow-only scratch buffer; one instance per (process, device)."""

    def __init__(self, device: int = 0, sample_rate: int = 16000):
        if not torch.cuda.is_available():
            raise SsrError("ssr_b200 augmentation needs a CUDA device; there is no CPU fallback")
        self._lib = _lib.load()
        self.device = int(device)
        self.sample_rate = int(sample_rate)
        self._work = None
        self._pin = None

    def _c_ops(self, ops: Sequence[AugOp]):
        arr = (_lib.AugOp * len(ops))()
        for i, op in enumerate(ops):
            arr[i] = _lib.AugOp(kind=_KIND_CODE[op.kind],
                                new_rate=int(op.n_steps if op.kind == "pitch" else op.new_rate),
                                factor=float(op.factor), reserved=0, seed=int(op.seed) & 0xFFFFFFFFFFFFFFFF)
        return arr

    def run_device(self, audio: torch.Tensor, n_samples, ops: Sequence[AugOp], noise: torch.Tensor | None = None,
                   out: torch.Tensor | None = None, stream=None) -> tuple[torch.Tensor, np.ndarray]:
        """audio: CUDA float32 [B, ld]; n_samples: host ints [B]; ops: one per clip; noise (optional): CUDA float32
        [B, >= ld] standard normals for the noise kind. Returns (CUDA float32 [B, ld_out] zero-padded, n_out[B])."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.stride(1) == 1
        B = audio.shape[0]
        assert len(ops) == B
        n_in = np.ascontiguousarray(n_samples, dtype=np.int32)
        c_ops = self._c_ops(ops)
        n_in_p = n_in.ctypes.data_as(_lib.c_i32p)
        # a speed round trip returns n, n+1 or n+2 samples (two ceilings); the library checks the real lengths
        ld_out = (int(n_in.max(initial=0)) + 4 + 7) // 8 * 8
        if out is None:
            out = torch.empty((B, ld_out), dtype=torch.float32, device=audio.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.stride(1) == 1
        need = int(self._lib.ssr_augment_work_bytes(n_in_p, B, c_ops, self.sample_rate))
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=audio.device)
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.stride(1) == 1
            assert noise.shape[0] == B and noise.shape[1] >= int(n_in.max(initial=0))
        n_out = np.zeros(B, dtype=np.int32)
        err = C.create_string_buffer(256)
        st = torch.cuda.current_stream(audio.device) if stream is None else stream
        rc = self._lib.ssr_augment(audio.data_ptr(), row_pitch(audio), n_in_p, B, c_ops, self.sample_rate,
                                   None if noise is None else noise.data_ptr(),
                                   0 if noise is None else row_pitch(noise), self._work.data_ptr(),
                                   self._work.numel(), out.data_ptr(), row_pitch(out),
                                   n_out.ctypes.data_as(_lib.c_i32p), st.cuda_stream, err, 256)
        if rc != 0:
            raise SsrError(err.value.decode(errors="replace"))
        return out, n_out

    def run(self, clips: Sequence, ops: Sequence[AugOp], noise: Sequence | None = None) -> list[np.ndarray]:
        """Host in, host out: list of 1-D clips -> list of augmented float32 clips."""
        if len(clips) == 0:
            return []
        audio, n = _stage(clips, self.device)
        nz = None
        if noise is not None:
            nz, _ = _stage([np.zeros(0, np.float32) if z is None else z for z in noise], self.device,
                           ld=audio.shape[1])
        out, n_out = self.run_device(audio, n, ops, nz)
        host = out.cpu().numpy()
        return [host[i, : n_out[i]].copy() for i in range(len(clips))]


def _stage(clips: Sequence, device: int, ld: int | None = None) -> tuple[torch.Tensor, np.ndarray]:
    arrs = [_as_f32_clip(c) for c in clips]
    n = np.array([a.size for a in arrs], dtype=np.int32)
    if ld is None:
        ld = int(max(8, (int(n.max(initial=0)) + 7) // 8 * 8))
    host = torch.zeros((len(arrs), ld), dtype=torch.float32).pin_memory()
    hn = host.numpy()
    for i, a in enumerate(arrs):
        hn[i, : a.size] = a
    return host.to(torch.device("cuda", device), non_blocking=True), n


_AUGMENTERS: dict = {}


def get_augmenter(device: int = 0, sample_rate: int = 16000) -> Augmenter:
    key = (int(device), int(sample_rate))
    if key not in _AUGMENTERS:
        _AUGMENTERS[key] = Augmenter(*key)
    return _AUGMENTERS[key]


# ---------------------------------------------------------------------------------------------- drop-in
def augment_audio(waveform, sample_rate=16000, augmentation_type="random", variant="model_training_1", device=0):
    """Same call and return as the reference's augment_audio: 1-D float32 numpy out; on any failure a warning is
    logged and the input comes back unchanged (REF/model_training_1.py:208-210)."""
    if isinstance(waveform, torch.Tensor):
        waveform = waveform.detach().cpu().numpy()
    x = np.ascontiguousarray(np.asarray(waveform, dtype=np.float32).reshape(-1))
    try:
        op = draw_op(augmentation_type, sample_rate, variant)
        noise = None
        if op.kind == "noise":  # the reference's generator and draw shape: randn_like of a [1, n] tensor
            noise = [torch.randn_like(torch.from_numpy(x).unsqueeze(0)).squeeze(0).numpy()]
        return get_augmenter(device, sample_rate).run([x], [op], noise)[0]
    except Exception as e:  # noqa: BLE001 - reference behaviour
        logger.warning(f"Augmentation failed: {e}. Returning original audio.")
        return x


# ---------------------------------------------------------------------------------------------- batched flow
def augment_and_extract(engine, clips: Sequence, augmentation_factor: int = 1, variant="model_training_1",
                        batch: int = 256, seed: int = 0, ops: Sequence[AugOp] | None = None,
                        sample_rate: int = 16000) -> tuple[np.ndarray, list[AugOp]]:
    """The inner loop of augment_minority_classes (REF/model_training_1.py:362-393) as batched device work: for every
    clip, `augmentation_factor` augmented versions are produced and embedded without leaving the GPU.

    Returns (pooled float32 [len(clips) * augmentation_factor, L+1, D] ordered clip-major like the reference's nested
    loops, the ops that were applied). `ops` may be given to replay a plan; otherwise decisions are drawn from
    python's `random` in the reference's order and noise comes from the device generator keyed by (seed, position)."""
    n_total = len(clips) * augmentation_factor
    if ops is None:
        ops = []
        for i in range(n_total):
            op = draw_op("random", sample_rate, variant)
            op.seed = (seed << 32) ^ i
            ops.append(op)
    assert len(ops) == n_total
    aug = get_augmenter(engine.device, sample_rate)
    out = np.zeros((n_total, engine.layers + 1, engine.hidden), np.float32)
    src = [clips[i // augmentation_factor] for i in range(n_total)]
    for lo in range(0, n_total, batch):
        hi = min(lo + batch, n_total)
        audio, n = _stage(src[lo:hi], engine.device)
        y, n_out = aug.run_device(audio, n, ops[lo:hi])
        out[lo:hi] = engine.pooled_device(y, n_out).cpu().numpy()
    return out, list(ops)
