"""The device GELU (csrc/common.cuh: gelu_fast) restated in numpy float32 with the coefficients read from the header.

Reference: erf-GELU, `torch.nn.functional.gelu` / HF "gelu" (HF/models/wavlm/modeling_wavlm.py:288-295,
HF/models/whisper/modeling_whisper.py:403-405). Guards the two properties the engine relies on: the absolute error
stays far below a bf16 ulp of the activations, and a zero input gives exactly zero (an all-zero clip must stay all-zero
through the conv + LayerNorm stack; tests/golden holds such a clip).
"""
import math
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stuttering-speech-representation_b200",
                   "csrc", "common.cuh")


def _coefficients():
    src = open(HDR).read()
    body = src[src.index("float gelu_fast(float x)"):]
    body = body[:body.index("return")]
    c = float(re.search(r"fmaf\(([0-9.eE+-]+)f, ax, 1\.0f\)", body).group(1))
    q3, q2 = re.search(r"float q = fmaf\((-?[0-9.eE+-]+)f, t, (-?[0-9.eE+-]+)f\);", body).groups()
    rest = re.findall(r"q = fmaf\(q, t, (-?[0-9.eE+-]+)f\);", body)
    assert len(rest) == 2, "gelu_fast is expected to be a cubic in t"
    return np.float32(c), [np.float32(v) for v in (q3, q2, rest[0], rest[1])]


def _fma(a, b, c):
    """a * b + c with one rounding to float32 (the product of two float32 is exact in float64)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def gelu_device(x):
    """float32 restatement of gelu_fast, instruction for instruction (MUFU results rounded to float32)."""
    c, (q3, q2, q1, q0) = _coefficients()
    f32 = np.float32
    x = x.astype(f32)
    ax = np.abs(x)
    t = (1.0 / _fma(c, ax, f32(1)).astype(np.float64)).astype(f32)           # rcp.approx
    q = _fma(q3, t, q2)
    q = _fma(q, t, q1)
    q = _fma(q, t, q0)
    arg = ((ax * f32(-0.5 * 1.4426950408889634)).astype(f32) * ax).astype(f32)
    ex = np.exp2(arg.astype(np.float64)).astype(f32)                          # ex2.approx
    return _fma(-q, ex, np.maximum(x, f32(0)))


def gelu_exact(x):
    x = x.astype(np.float64)
    return 0.5 * x * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))


def test_gelu_absolute_error_and_zero():
    x = np.linspace(-12.0, 12.0, 240001).astype(np.float32)
    err = np.abs(gelu_device(x).astype(np.float64) - gelu_exact(x))
    assert err.max() <= 1.2e-5, err.max()          # stated in common.cuh: 9.2e-6 (+ MUFU rounding on the device)
    z = gelu_device(np.zeros(4, np.float32))
    assert (z == 0).all() and not np.signbit(z).any(), z


def test_gelu_relative_error_for_small_inputs():
    # tiny random-init models run at activations of 1e-3 .. 1e-2: the slope at 0 must be right, not just the value
    x = np.concatenate([np.logspace(-3, 0, 301), -np.logspace(-3, 0, 301)]).astype(np.float32)
    rel = np.abs(gelu_device(x).astype(np.float64) - gelu_exact(x)) / np.abs(gelu_exact(x))
    assert rel.max() <= 5e-4, rel.max()


def test_gelu_tails():
    x = np.array([-40.0, -15.0, 15.0, 40.0, 1e4, -1e4], np.float32)
    y = gelu_device(x)
    assert np.isfinite(y).all()
    np.testing.assert_allclose(y, np.maximum(x, 0), atol=1e-6)
