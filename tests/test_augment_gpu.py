"""GPU tier for the augmentation row: the CUDA path (through the C ABI `ssr_augment`) against the
fixture produced by the reference's own augment_audio, against the numpy oracle, and size-independent properties at
the BASELINE batch size."""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "augment.npz")
SUB = 4
# Stated tolerance of the resampling kinds: fp32 tap products summed in a different order than torch's conv1d.
RESAMPLE_TOL = 2e-6


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def aug():
    from ssr_b200 import augment

    return augment.get_augmenter(0)


def test_drop_in_vs_reference_golden(gold):
    """augment_audio (same signature, same RNG streams) against the reference's outputs."""
    import torch

    from ssr_b200 import augment, synth

    clips = synth.aug_clips()
    n_speed = 0
    for name in (str(n) for n in gold["names"]):
        variant, kind, seed = name.split("/")
        seed = int(seed)
        ci = seed % 3 if kind != "none" else 1
        random.seed(seed)
        drawn = augment.draw_op(kind, 16000, variant).kind
        random.seed(seed)
        torch.manual_seed(seed)
        out = augment.augment_audio(clips[ci].copy(), augmentation_type=kind, variant=variant)
        assert out.dtype == np.float32 and out.ndim == 1 and out.shape[0] == int(gold[name + "/len"]), name
        err = float(np.abs(out[::SUB] - gold[name + "/sub"]).max())
        if drawn == "speed":
            n_speed += 1
            assert err <= RESAMPLE_TOL, (name, err)
        else:
            assert err == 0.0, (name, err)  # noise (replayed torch normals), volume, none: bit-exact
        assert abs(float(out.astype(np.float64).sum()) - float(gold[name + "/sum"])) <= 2e-3
    assert n_speed >= 6


@pytest.mark.parametrize("nr", [16001, 15200, 16777])
def test_resampler_vs_torchaudio_golden(gold, aug, nr):
    from ssr_b200 import synth
    from ssr_b200.augment import AugOp

    x = synth.aug_clips()[0][:24000]
    out = aug.run([x], [AugOp("speed", new_rate=nr)])[0]
    assert out.shape[0] == int(gold[f"resample/{nr}/lens"][1])
    ref = np.clip(gold[f"resample/{nr}/out_sub"], -1, 1)
    assert np.abs(out[::SUB] - ref).max() <= RESAMPLE_TOL


def test_mixed_ragged_batch_vs_oracle(aug):
    """One batch mixing every kind, ragged lengths (incl. empty and one-sample clips), random perturbed rates."""
    from oracle import augment_oracle as ao
    from ssr_b200.augment import AugOp

    rng = np.random.default_rng(5)
    lens = [48000, 1, 0, 16001, 7, 31999, 48000, 12345, 400, 47999, 20000, 33333]
    clips = [(rng.standard_normal(n) * 0.3).astype(np.float32) for n in lens]
    clips[6] = (clips[6] * 4).astype(np.float32)  # clamp bites
    rates = [15200, 16001, 15999, 16799, 15201, 16000, 16640, 15873, 16384, 14401, 17599, 16123]
    ops, noise = [], []
    for i, n in enumerate(lens):
        k = ("speed", "noise", "volume", "none")[i % 4] if i >= 4 else "speed"
        ops.append(AugOp(k, new_rate=rates[i], factor=[0.0, 0.004, 1.1, 0.0][i % 4] if i >= 4 else 0.0))
        noise.append(rng.standard_normal(n).astype(np.float32) if k == "noise" else None)
    got = aug.run(clips, ops, noise)
    for i, (x, op) in enumerate(zip(clips, ops)):
        want = ao.apply(x, op.kind, {"new_rate": op.new_rate, "factor": op.factor}, 16000, noise[i])
        assert got[i].shape == want.shape, (i, op)
        if want.size:
            err = float(np.abs(got[i] - want).max())
            assert err <= (RESAMPLE_TOL if op.kind == "speed" else 0.0), (i, op, err)


def test_output_is_zero_padded_and_lengths_reported(aug):
    import torch

    from ssr_b200.augment import AugOp

    x = torch.randn(3, 4000, device="cuda") * 0.1
    out, n_out = aug.run_device(x, [4000, 1000, 2500], [AugOp("speed", new_rate=16501), AugOp("volume", factor=0.9),
                                                        AugOp("none")])
    assert list(n_out) == [4001, 1000, 2500]
    o = out.cpu().numpy()
    for b in range(3):
        assert not o[b, n_out[b]:].any()
    np.testing.assert_array_equal(o[2, :2500], x[2, :2500].cpu().numpy())


def test_device_noise_statistics_and_determinism(aug):
    import torch

    from ssr_b200.augment import AugOp

    B, n = 8, 48000
    x = torch.zeros(B, n, device="cuda")
    ops = [AugOp("noise", factor=0.01, seed=100 + b) for b in range(B)]
    a, _ = aug.run_device(x, [n] * B, ops)
    a = a.cpu().numpy()[:, :n].astype(np.float64) / 0.01
    assert abs(a.mean()) < 0.01 and abs(a.std() - 1.0) < 0.01
    assert abs((a ** 3).mean()) < 0.03 and abs((a ** 4).mean() - 3.0) < 0.08   # skewness / kurtosis of N(0,1)
    assert np.abs(np.corrcoef(a[0], a[1])[0, 1]) < 0.02                         # different seeds are independent
    assert np.abs(np.corrcoef(a[0, :-1], a[0, 1:])[0, 1]) < 0.02                # white
    # keyed by the op's seed only: batch position and batch size do not matter
    perm = [3, 0, 7, 1]
    b, _ = aug.run_device(x[:4], [n] * 4, [ops[i] for i in perm])
    np.testing.assert_array_equal(b.cpu().numpy()[:, :n].astype(np.float64) / 0.01, a[perm])


def test_properties_at_baseline_batch(aug):
    """B = 256 x 3 s (BASELINE configs[1] shape): lengths, identity rate, linearity of the round trip, clamp bound."""
    import torch

    from ssr_b200.augment import AugOp, out_length

    B, n = 256, 48000
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(B, n, device="cuda", generator=g) * 0.1
    rng = random.Random(9)
    ops = [AugOp("speed", new_rate=int(16000 * rng.uniform(0.9, 1.1))) for _ in range(B)]
    ops[0] = AugOp("speed", new_rate=16000)  # Resample(sr, sr) is the identity
    y, n_out = aug.run_device(x, [n] * B, ops)
    assert [int(v) for v in n_out] == [out_length(op, n) for op in ops]
    assert all(abs(int(v) - n) <= 2 for v in n_out)
    torch.testing.assert_close(y[0, :n], x[0], rtol=0, atol=0)
    assert float(y.abs().max()) <= 1.0
    # linearity: resample(0.5 x) == 0.5 resample(x) exactly (power-of-two scale commutes with every rounding)
    y2, _ = aug.run_device(x * 0.5, [n] * B, ops)
    torch.testing.assert_close(y2, y * 0.5, rtol=0, atol=0)
    # a band-limited round trip returns (nearly) the input away from the edges: slow sine, 300 Hz
    t = torch.arange(n, device="cuda", dtype=torch.float64) / 16000.0
    s = (0.5 * torch.sin(2 * np.pi * 300.0 * t)).float().repeat(B, 1)
    z, nz = aug.run_device(s, [n] * B, ops)
    mid = slice(200, n - 200)
    assert float((z[:, mid] - s[:, mid]).abs().max()) < 2e-3


def test_pitch_vs_reference_golden_and_oracle(gold):
    """The pitch kind (REF/model_training_01.py:174-178) through the drop-in, same seeds as the reference run.
    Tolerances: tests/test_augment_cpu.py PITCH_TOL (tight on broadband clips, statistical on the tonal chirp)."""
    import torch

    from oracle import augment_oracle as ao
    from ssr_b200 import augment, synth
    from test_augment_cpu import PITCH_TOL, pitch_errors  # tests/ is on sys.path (pytest rootdir-less layout)

    clips = synth.aug_clips()
    for name in (str(n) for n in gold["pitch_names"]):
        _, ci, n_steps = name.split("/")
        ci, n_steps = int(ci), int(n_steps)
        seed = int(gold[name + "/seed"])
        random.seed(seed)
        torch.manual_seed(seed)
        out = augment.augment_audio(clips[ci].copy(), augmentation_type="pitch", variant="model_training_01")
        assert out.dtype == np.float32 and out.shape[0] == int(gold[name + "/len"]), name
        assert not np.array_equal(out, clips[ci]), "augmentation fell back to the input"
        mx, rel = pitch_errors(out, gold, name)
        tol_max, tol_rel = PITCH_TOL[ci]
        assert rel <= tol_rel and (tol_max is None or mx <= tol_max), (name, mx, rel)
        if ci == 2:  # and against the oracle on the short clip
            want = ao.apply(clips[ci], "pitch", {"n_steps": n_steps})
            assert np.abs(out - want).max() <= 1e-3


def test_pitch_batch_properties(aug):
    """Mixed batch with the pitch kind: n_steps = 0 is the identity, short clips are rejected like torch.stft rejects
    them, the pitch really moves (spectral centroid of a harmonic tone scales by 2^(n/12)), zero padding kept."""
    import torch

    from ssr_b200.augment import AugOp
    from ssr_b200.engine import SsrError

    n = 32000
    t = np.arange(n) / 16000.0
    tone = (0.3 * np.sin(2 * np.pi * 440.0 * t) + 0.05 * np.random.default_rng(0).standard_normal(n)).astype(np.float32)
    outs = aug.run([tone, tone, tone, tone[:20000]],
                   [AugOp("pitch", n_steps=2), AugOp("pitch", n_steps=-2), AugOp("pitch", n_steps=0),
                    AugOp("volume", factor=0.5)])
    assert [o.shape[0] for o in outs] == [n, n, n, 20000]
    np.testing.assert_array_equal(outs[2], tone)

    def peak_hz(x):
        spec = np.abs(np.fft.rfft(x[4000:-4000] * np.hanning(len(x) - 8000)))
        return np.argmax(spec) * 16000.0 / (len(x) - 8000)

    assert abs(peak_hz(outs[0]) / 440.0 - 2 ** (2 / 12)) < 0.01
    assert abs(peak_hz(outs[1]) / 440.0 - 2 ** (-2 / 12)) < 0.01
    with pytest.raises(SsrError):
        aug.run([tone[:200]], [AugOp("pitch", n_steps=1)])
    # the drop-in turns that failure into "return the input" like the reference's except branch
    from ssr_b200 import augment
    random.seed(0)  # randint(-2, 2) -> 1
    short = tone[:200].copy()
    np.testing.assert_array_equal(augment.augment_audio(short, augmentation_type="pitch", variant="model_training_01"),
                                  short)


def test_augment_and_extract_matches_separate_steps(aug):
    """Batched augment -> encoder on the device equals oracle-augmented clips pushed through the same engine."""
    from oracle import augment_oracle as ao
    from ssr_b200 import WavLMEngine, augment, synth

    model, fe = synth.build_wavlm("tiny_stable")
    eng = WavLMEngine.from_hf(model, fe)
    clips = synth.aug_clips()
    random.seed(11)
    pooled, ops = augment.augment_and_extract(eng, clips, augmentation_factor=3, batch=4, seed=1)
    assert pooled.shape == (9, eng.layers + 1, eng.hidden) and len(ops) == 9
    kinds = {op.kind for op in ops}
    assert "speed" in kinds
    for i, op in enumerate(ops):
        if op.kind == "noise":
            continue  # device generator: no host replay
        x = ao.apply(clips[i // 3], op.kind, {"new_rate": op.new_rate, "factor": op.factor})
        want = eng.pooled([x])[0]
        rel = np.abs(pooled[i] - want).max() / np.abs(want).max()
        assert rel < 2e-3, (i, op, rel)
