"""TEST INFRASTRUCTURE — CPU restatement of the reference's waveform augmentation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.

Follows
    augment_audio   /root/reference/model_training_1.py:166-213   (variant "model_training_1")
    augment_audio   /root/reference/model_training_01.py:140-192  (variant "model_training_01", incl. the pitch kind)
whose resampling arithmetic lives in the un-vendored third-party dependency torchaudio (2.11 here; unpinned by the
reference): torchaudio/functional/functional.py `_get_sinc_resample_kernel` and `_apply_sinc_resample_kernel` with
the transform's defaults (sinc_interp_hann, lowpass_filter_width=6, rolloff=0.99, kernel built in float64 and cast to
float32, conv1d in float32).

Pinned by tests/golden/augment.npz, produced by the reference's own augment_audio (tools/make_golden_aug.py).

One deliberate difference in evaluation (not in result): torchaudio materialises the whole [new/gcd, 2*width+orig/gcd]
filter bank; taps with |t| >= 6 carry the window value cos(pi/2)^2 ~ 3.7e-33 and are skipped here.
"""
from __future__ import annotations

import math
import random

import numpy as np

LOWPASS_WIDTH = 6
ROLLOFF = 0.99

VARIANTS = {
    # kinds in the reference's random.choice order; (lo, hi) of the random.uniform draws
    "model_training_1": {"kinds": ["speed", "noise", "volume", "none"], "speed": (0.95, 1.05),
                         "noise": (0.001, 0.005), "volume": (0.9, 1.1)},
    "model_training_01": {"kinds": ["speed", "noise", "pitch", "volume"], "speed": (0.9, 1.1),
                          "noise": (0.005, 0.02), "volume": (0.8, 1.2)},
}


def resample_length(n: int, orig: int, new: int) -> int:
    """`torch.ceil(torch.as_tensor(new_freq * length / orig_freq)).long()`: python-float quotient, float32 tensor."""
    if orig == new:
        return n
    g = math.gcd(orig, new)
    o, nw = orig // g, new // g
    return int(np.ceil(np.float32(nw * n / o)))


def sinc_resample(x: np.ndarray, orig: int, new: int) -> np.ndarray:
    """torchaudio.transforms.Resample(orig, new)(x) for a 1-D float32 signal."""
    x = np.asarray(x, np.float32)
    if orig == new:
        return x
    g = math.gcd(int(orig), int(new))
    o, nw = int(orig) // g, int(new) // g
    base = min(o, nw) * ROLLOFF
    width = math.ceil(LOWPASS_WIDTH * o / base)
    scale = base / o
    length = x.shape[0]
    n_out = resample_length(length, orig, new)
    # padded signal exactly as `_apply_sinc_resample_kernel` pads it
    xp = np.concatenate([np.zeros(width, np.float32), x, np.zeros(width + o, np.float32)])
    n = np.arange(n_out, dtype=np.int64)
    q, p = n // nw, n % nw
    # taps j of the reference kernel row p (j = 0 .. 2*width+o-1); only those around the centre are non-negligible
    half = int(math.ceil(LOWPASS_WIDTH * o / base)) + 1
    centre = width + (p * o) // nw
    j = centre[:, None] + np.arange(-half, half + 2, dtype=np.int64)[None, :]
    ok = (j >= 0) & (j < 2 * width + o)
    jc = np.clip(j, 0, 2 * width + o - 1)
    idx = (jc - width).astype(np.float64) / o                      # torch.arange(-width, width+orig)/orig
    # `torch.arange(0, -new_freq, -1, dtype=None) / new_freq`: an int64 tensor divided by an int is a FLOAT32 true
    # division; only then is it promoted to float64 by the addition. The rounded phase is part of the reference's
    # observable behaviour (up to ~5e-5 on a unit-scale signal for co-prime rates), so it is restated, not "fixed".
    phase = ((-p).astype(np.float32) / np.float32(nw)).astype(np.float64)
    t = phase[:, None] + idx                                       # arange(0,-new,-1)/new + idx
    t = t * base
    t = np.clip(t, -LOWPASS_WIDTH, LOWPASS_WIDTH)
    window = np.cos(t * math.pi / LOWPASS_WIDTH / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = (k * (window * scale)).astype(np.float32)
    k = np.where(ok, k, np.float32(0))
    src = q[:, None] * o + jc
    src_ok = src < xp.shape[0]
    vals = np.where(src_ok, xp[np.minimum(src, xp.shape[0] - 1)], np.float32(0))
    out = (vals.astype(np.float64) * k.astype(np.float64)).sum(1)
    return out.astype(np.float32)


def draw(augmentation_type="random", sample_rate=16000, variant="model_training_1", rng=random):
    """The reference's host-side decisions, in its draw order. Returns (kind, params)."""
    v = VARIANTS[variant]
    kind = augmentation_type
    if kind == "random":
        kind = rng.choice(v["kinds"])
    if kind == "speed":
        f = rng.uniform(*v["speed"])
        return kind, {"new_rate": int(sample_rate * f), "speed_factor": f}
    if kind == "noise":
        return kind, {"factor": rng.uniform(*v["noise"])}
    if kind == "volume":
        return kind, {"factor": rng.uniform(*v["volume"])}
    if kind == "pitch":
        return kind, {"n_steps": rng.randint(-2, 2)}
    return "none", {}


def apply(x: np.ndarray, kind: str, params: dict, sample_rate=16000, noise=None) -> np.ndarray:
    """Deterministic part of augment_audio given the drawn decision (and, for noise, the standard normals)."""
    x = np.asarray(x, np.float32)
    if kind == "speed":
        nr = params["new_rate"]
        x = sinc_resample(sinc_resample(x, sample_rate, nr), nr, sample_rate)
    elif kind == "noise":
        z = np.asarray(noise, np.float32)
        x = x + z * np.float32(params["factor"])
    elif kind == "volume":
        x = x * np.float32(params["factor"])
    elif kind == "pitch":
        x = pitch_shift(x, params["n_steps"], sample_rate)
    elif kind != "none":
        raise NotImplementedError(kind)
    return np.clip(x, np.float32(-1.0), np.float32(1.0)).astype(np.float32)


def augment_audio(waveform, sample_rate=16000, augmentation_type="random", variant="model_training_1"):
    """Full restatement including the RNG streams: python `random` for the decision, torch's CPU generator for the
    noise (torch.randn_like), so that seeding both like the reference reproduces its output."""
    import torch

    x = np.asarray(waveform, np.float32)
    kind, params = draw(augmentation_type, sample_rate, variant)
    noise = None
    if kind == "noise":
        noise = torch.randn_like(torch.from_numpy(x).unsqueeze(0)).squeeze(0).numpy()
    return apply(x, kind, params, sample_rate, noise)


# ---------------------------------------------------------------------------------------------- pitch shift
# torchaudio.transforms.PitchShift(sample_rate, n_steps) as used by REF/model_training_01.py:174-178:
# functional `_stretch_waveform` (STFT 512 / hop 128 / periodic hann, phase vocoder, inverse STFT) followed by the
# sinc resampler whose filter bank — unlike transforms.Resample — is built ENTIRELY IN FLOAT32 (the transform passes
# dtype=input.dtype to `_get_sinc_resample_kernel`), then crop / zero-pad to the input length.
N_FFT, HOP = 512, 128


def pitch_rate(n_steps: int) -> float:
    return 2.0 ** (-float(n_steps) / 12)


def _hann_periodic(n: int) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)


def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace in float32: symmetric evaluation around the midpoint."""
    start, end = np.float32(start), np.float32(end)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps)
    lo = (start + step * i.astype(np.float32)).astype(np.float32)
    hi = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < steps // 2, lo, hi).astype(np.float32)


def stft_512(x: np.ndarray) -> np.ndarray:
    """torch.stft(x, 512, 128, 512, hann, center=True, reflect, onesided) -> complex64 [frames, 257]."""
    x = np.asarray(x, np.float32)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = 1 + x.shape[0] // HOP
    idx = np.arange(n_frames)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = xp[idx] * _hann_periodic(N_FFT)[None, :]
    return np.fft.rfft(frames.astype(np.float64), axis=1).astype(np.complex64)


def phase_vocoder(spec: np.ndarray, rate: float) -> np.ndarray:
    """torchaudio.functional.phase_vocoder on [frames, 257] complex64 (frame-major here)."""
    F = spec.shape[0]
    n_out = int(np.ceil(F / rate))
    ts = (rate * np.arange(n_out, dtype=np.float64)).astype(np.float32)      # torch.arange(0, F, rate, float32)
    alphas = np.fmod(ts, np.float32(1.0)).astype(np.float32)
    pa = _linspace_f32(0.0, np.pi * HOP, spec.shape[1])                       # [257]
    phase_0 = np.angle(spec[:1]).astype(np.float32)                           # [1, 257]
    sp = np.concatenate([spec, np.zeros((2, spec.shape[1]), np.complex64)])
    i0 = ts.astype(np.int64)
    i1 = (ts + np.float32(1.0)).astype(np.int64)
    s0, s1 = sp[i0], sp[i1]
    a0 = np.arctan2(s0.imag, s0.real).astype(np.float32)
    a1 = np.arctan2(s1.imag, s1.real).astype(np.float32)
    n0 = np.hypot(s0.real, s0.imag).astype(np.float32)
    n1 = np.hypot(s1.real, s1.imag).astype(np.float32)
    two_pi = np.float32(2 * np.pi)
    ph = (a1 - a0 - pa[None, :]).astype(np.float32)
    ph = (ph - two_pi * np.round(ph / two_pi)).astype(np.float32)
    ph = (ph + pa[None, :]).astype(np.float32)
    ph = np.concatenate([phase_0, ph[:-1]])
    acc = np.cumsum(ph.astype(np.float64), axis=0).astype(np.float32)         # CPU cumsum accumulates in double
    mag = (alphas[:, None] * n1 + (np.float32(1.0) - alphas[:, None]) * n0).astype(np.float32)
    return (mag * np.cos(acc) + 1j * (mag * np.sin(acc))).astype(np.complex64)


def istft_512(spec: np.ndarray, length: int) -> np.ndarray:
    """torch.istft(spec, 512, 128, 512, hann, center=True, length=length) for frame-major [frames, 257]."""
    F = spec.shape[0]
    w = _hann_periodic(N_FFT).astype(np.float64)
    frames = np.fft.irfft(spec.astype(np.complex128), n=N_FFT, axis=1) * w[None, :]
    total = N_FFT + HOP * (F - 1)
    y = np.zeros(total)
    env = np.zeros(total)
    for f in range(F):
        y[f * HOP:f * HOP + N_FFT] += frames[f]
        env[f * HOP:f * HOP + N_FFT] += w * w
    start = N_FFT // 2
    end = min(start + length, total)
    out = (y[start:end] / env[start:end]).astype(np.float32)
    if out.shape[0] < length:
        out = np.concatenate([out, np.zeros(length - out.shape[0], np.float32)])
    return out


def sinc_resample_f32(x: np.ndarray, orig: int, new: int) -> np.ndarray:
    """The resampler with torchaudio's filter bank evaluated in float32 throughout (dtype=float32 path)."""
    x = np.asarray(x, np.float32)
    if orig == new:
        return x
    f32 = np.float32
    g = math.gcd(int(orig), int(new))
    o, nw = int(orig) // g, int(new) // g
    base = min(o, nw) * ROLLOFF
    width = math.ceil(LOWPASS_WIDTH * o / base)
    scale = f32(base / o)
    length = x.shape[0]
    n_out = resample_length(length, orig, new)
    xp = np.concatenate([np.zeros(width, f32), x, np.zeros(width + o, f32)])
    n = np.arange(n_out, dtype=np.int64)
    q, p = n // nw, n % nw
    half = width + 2
    centre = width + (p * o) // nw
    j = centre[:, None] + np.arange(-half, half + 2, dtype=np.int64)[None, :]
    ok = (j >= 0) & (j < 2 * width + o)
    jc = np.clip(j, 0, 2 * width + o - 1)
    idx = ((jc - width).astype(f32) / f32(o)).astype(f32)
    phase = ((-p).astype(f32) / f32(nw)).astype(f32)
    t = (phase[:, None] + idx).astype(f32)
    t = (t * f32(base)).astype(f32)
    t = np.clip(t, f32(-LOWPASS_WIDTH), f32(LOWPASS_WIDTH))
    window = (np.cos(((t * f32(math.pi)).astype(f32) / f32(LOWPASS_WIDTH)).astype(f32) / f32(2)).astype(f32) ** 2).astype(f32)
    t = (t * f32(math.pi)).astype(f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, f32(1.0), (np.sin(t).astype(f32) / t).astype(f32)).astype(f32)
    k = (k * (window * scale).astype(f32)).astype(f32)
    k = np.where(ok, k, f32(0))
    src = q[:, None] * o + jc
    vals = np.where(src < xp.shape[0], xp[np.minimum(src, xp.shape[0] - 1)], f32(0))
    return (vals.astype(np.float64) * k.astype(np.float64)).sum(1).astype(np.float32)


def pitch_shift(x: np.ndarray, n_steps: int, sample_rate=16000) -> np.ndarray:
    x = np.asarray(x, np.float32)
    if n_steps == 0:
        return x
    rate = pitch_rate(n_steps)
    spec = phase_vocoder(stft_512(x), rate)
    y = istft_512(spec, int(round(x.shape[0] / rate)))
    orig = int(sample_rate / rate)
    z = sinc_resample_f32(y, orig, sample_rate) if orig != sample_rate else y
    if z.shape[0] > x.shape[0]:
        return z[: x.shape[0]]
    return np.concatenate([z, np.zeros(x.shape[0] - z.shape[0], np.float32)])
