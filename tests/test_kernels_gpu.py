"""Kernel-level parity (through the C ABI) against plain PyTorch fp32 references of the same op."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


DEFAULT_VARIANT = 1  # csrc/attention_tc.cu g_attention_variant


def _lib():
    from ssr_b200 import _lib as L

    return L.load()


def _err():
    return C.create_string_buffer(512)


def _ptr(t):
    return None if t is None else t.data_ptr()


def gemm(A, lda, a_rows, W, M, N, K, bias=None, act=0, resid=None, want_f32=True, want_bf16=False, simt=0):
    lib = _lib()
    o32 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32) if want_f32 else None
    o16 = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16) if want_bf16 else None
    e = _err()
    rc = lib.ssr_gemm_bf16(0, A.data_ptr(), lda, a_rows, W.data_ptr(), M, N, K, _ptr(bias), act, _ptr(resid),
                           _ptr(o32), _ptr(o16), simt, None, e, 512)
    torch.cuda.synchronize()
    assert rc == 0, e.value.decode()
    return o32, o16


def ref_gemm(A2d, W, bias, act, resid):
    y = A2d.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    if act:
        y = torch.nn.functional.gelu(y)
    if resid is not None:
        y = y + resid
    return y


SHAPES = [
    (128, 256, 64), (128, 256, 128), (256, 512, 256), (300, 1024, 512), (149 * 3, 1024, 1024),
    (1000, 768, 3072), (128, 128, 64), (200, 384, 192), (130, 64, 128), (512, 1280, 5120), (777, 3840, 1280),
]


@pytest.mark.parametrize("simt", [0, 1], ids=["tc", "simt"])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_plain(M, N, K, simt):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    o32, _ = gemm(A, K, M, W, M, N, K, simt=simt)
    ref = ref_gemm(A, W, None, 0, None)
    err = (o32 - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * max(scale, 1.0), f"max err {err} (ref max {scale}); first rows {o32[:2, :4]} vs {ref[:2, :4]}"


@pytest.mark.parametrize("simt", [0, 1], ids=["tc", "simt"])
def test_gemm_epilogue_bias_gelu_resid_bf16(simt):
    M, N, K = 333, 512, 320
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    o32, o16 = gemm(A, K, M, W, M, N, K, bias=bias, act=1, resid=resid, want_bf16=True, simt=simt)
    ref = ref_gemm(A, W, bias, 1, resid)
    assert (o32 - ref).abs().max().item() < 5e-3
    assert (o16.float() - ref).abs().max().item() < 5e-2


@pytest.mark.parametrize("simt", [0, 1], ids=["tc", "simt"])
@pytest.mark.parametrize("C_,k,s", [(512, 3, 2), (512, 2, 2), (80, 3, 1)])
def test_gemm_conv_view(C_, k, s, simt):
    """Conv1d over a channels-last signal as a GEMM whose A rows overlap (row pitch = stride * C < K = k * C)."""
    T_in, Co = 601, 256
    T_out = (T_in - k) // s + 1
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(T_in + 4, C_, device="cuda", generator=g).bfloat16()  # a few spare rows behind the signal
    w = (torch.randn(Co, C_, k, device="cuda", generator=g) * 0.05).bfloat16()
    Wg = w.permute(0, 2, 1).contiguous().reshape(Co, k * C_)
    o32, _ = gemm(x, s * C_, T_out, Wg, T_out, Co, k * C_, simt=simt)
    ref = torch.nn.functional.conv1d(x[:T_in].float().t().unsqueeze(0), w.float(), stride=s)[0].t()
    err = (o32 - ref).abs().max().item()
    assert err < 5e-3 * max(1.0, ref.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("simt", [0, 1], ids=["tc", "simt"])
@pytest.mark.parametrize("slot,lens", [(149, None), (150, [149, 150, 100, 33]), (64, [64, 1, 17, 64])])
def test_gemm_fused_pool(slot, lens, simt):
    lib = _lib()
    B = 4
    M, N, K = B * slot, 256, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
    resid = torch.randn(M, N, device="cuda", generator=g)
    lens_t = torch.tensor(lens if lens is not None else [slot] * B, device="cuda", dtype=torch.int32)
    out = torch.zeros(M, N, device="cuda")
    part = torch.full(((M + 31) // 32 * 2 * N,), float("nan"), device="cuda")
    pooled = torch.full((B, 3, N), float("nan"), device="cuda")
    e = _err()
    rc = lib.ssr_gemm_bf16_pool(0, A.data_ptr(), K, W.data_ptr(), M, N, K, None, 0, resid.data_ptr(), out.data_ptr(),
                                slot, lens_t.data_ptr(), B, part.data_ptr(), pooled[:, 1].data_ptr(), 3 * N, simt,
                                None, e, 512)
    torch.cuda.synchronize()
    assert rc == 0, e.value.decode()
    ref = ref_gemm(A, W, None, 0, resid).view(B, slot, N)
    for b in range(B):
        L = int(lens_t[b])
        want = ref[b, :L].mean(0)
        got = pooled[b, 1]
        assert (got - want).abs().max().item() < 2e-3, (b, (got - want).abs().max().item())


@pytest.mark.parametrize("D", [512, 768, 1024, 1280])
@pytest.mark.parametrize("bf16_in,gelu", [(False, 0), (True, 1)])
def test_layernorm(D, bf16_in, gelu):
    lib = _lib()
    rows = 1003
    g = torch.Generator(device="cuda").manual_seed(D)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 0.5
    if bf16_in:
        x = x.bfloat16()
    gam = torch.randn(D, device="cuda", generator=g)
    bet = torch.randn(D, device="cuda", generator=g)
    o32 = torch.empty(rows, D, device="cuda")
    o16 = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    e = _err()
    rc = lib.ssr_layernorm(None if bf16_in else x.data_ptr(), x.data_ptr() if bf16_in else None, rows, D,
                           gam.data_ptr(), bet.data_ptr(), gelu, o32.data_ptr(), o16.data_ptr(), None, e, 512)
    torch.cuda.synchronize()
    assert rc == 0, e.value.decode()
    ref = torch.nn.functional.layer_norm(x.float(), (D,), gam, bet, 1e-5)
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert (o32 - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert (o16.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


def _attn_ref(qkv, B, slot, H, lens, gate, relbias, rel_center):
    D = H * 64
    x = qkv.float().view(B, slot, 3, H, 64)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)  # [B,H,T,64]
    s = q @ k.transpose(-1, -2)
    if gate is not None:
        i = torch.arange(slot, device="cuda")
        rel = i[None, :] - i[:, None] + rel_center  # [i, j] -> j - i
        bias = relbias[:, rel]  # [H, T, T]
        s = s + gate.view(B, slot, H).permute(0, 2, 1).unsqueeze(-1) * bias.unsqueeze(0)
    j = torch.arange(slot, device="cuda")
    mask = j[None, :] >= lens[:, None].long()
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = p @ v
    return o.transpose(1, 2).reshape(B * slot, D)


@pytest.mark.parametrize("impl", [0, 1], ids=["tc", "mma"])
@pytest.mark.parametrize("slot,H,lens,bias", [
    (149, 16, None, True), (150, 12, [149, 77, 150], True), (200, 4, [200, 64, 65], False), (1500, 2, None, False),
    (31, 3, [31, 5, 1], True), (128, 2, [128, 127, 1], True), (257, 2, [257, 256, 129], True),
    (1500, 2, [1500, 1400, 300], True), (160, 4, [160, 129, 33], True), (64, 2, [64, 32, 1], False),
    (160, 3, [159, 97, 160], False), (161, 2, [161, 160, 2], True),
])
def test_attention(slot, H, lens, bias, impl):
    lib = _lib()
    B = 3
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(slot + H)
    qkv = torch.randn(B * slot, 3 * D, device="cuda", generator=g)
    qkv[:, :D] *= 0.125 * 2.0
    qkv = qkv.bfloat16()
    lens_t = torch.tensor(lens if lens is not None else [slot] * B, device="cuda", dtype=torch.int32)
    R = 2048
    gate = relb = None
    if bias:
        gate = torch.rand(B * slot, H, device="cuda", generator=g) * 2 + 0.5
        relb = torch.randn(H, 2 * R - 1, device="cuda", generator=g)
    ref = _attn_ref(qkv, B, slot, H, lens_t, gate, relb, R - 1).view(B, slot, D)
    if impl == 0 and 128 < slot <= 256:
        # two-tile clips are walked in the paired item order by default; the query-tile-major order must agree
        assert lib.ssr_tuning_set(b"attention_paired", 0) == 0
        out = torch.zeros(B * slot, D, device="cuda", dtype=torch.bfloat16)
        e = _err()
        rc = lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens_t.data_ptr(), _ptr(gate), _ptr(relb),
                               2 * R - 1, R - 1, 0, None, e, 512)
        torch.cuda.synchronize()
        lib.ssr_tuning_set(b"attention_paired", 1)
        assert rc == 0, e.value.decode()
        got = out.float().view(B, slot, D)
        for b in range(B):
            L = int(lens_t[b])
            assert (got[b, :L] - ref[b, :L]).abs().max().item() < 2e-2, f"query-tile-major order, clip {b}"
    # every variant of the tcgen05 kernel (S prefetch from TMEM, polynomial exp2 share) must hold the same tolerance
    # every arithmetic variant of the tcgen05 kernel (packed pairs, polynomial exp2 share) must hold the same tolerance
    for variant in ([0, 1, 2, 3] if impl == 0 else [DEFAULT_VARIANT]):
        assert lib.ssr_tuning_set(b"attention_variant", variant) == 0
        out = torch.zeros(B * slot, D, device="cuda", dtype=torch.bfloat16)
        e = _err()
        rc = lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens_t.data_ptr(), _ptr(gate), _ptr(relb),
                               2 * R - 1, R - 1, impl, None, e, 512)
        torch.cuda.synchronize()
        lib.ssr_tuning_set(b"attention_variant", DEFAULT_VARIANT)
        assert rc == 0, e.value.decode()
        got = out.float().view(B, slot, D)
        for b in range(B):
            L = int(lens_t[b])
            err = (got[b, :L] - ref[b, :L]).abs().max().item()
            assert err < 2e-2, f"variant {variant}, clip {b}: max err {err}"


@pytest.mark.parametrize("B,slot,H,bias", [(40, 640, 4, False), (24, 385, 6, True), (64, 300, 5, True)])
def test_attention_item_orders_bit_identical(B, slot, H, bias):
    """More items than resident CTAs and three or more query tiles per clip: the persistent CTAs walk the item list
    either query-tile-major or grouped (all tiles of a (clip, head) on neighbouring CTAs at the same time, knob
    "attention_grouped"). The work per item is the same, so the outputs must be bit-identical, and right."""
    lib = _lib()
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(B + slot)
    qkv = torch.randn(B * slot, 3 * D, device="cuda", generator=g)
    qkv[:, :D] *= 0.25
    qkv = qkv.bfloat16()
    lens = [max(1, slot - 37 * i % slot) for i in range(B)]
    lens[0], lens[-1] = slot, 1
    lens_t = torch.tensor(lens, device="cuda", dtype=torch.int32)
    R = 2048
    gate = relb = None
    if bias:
        gate = torch.rand(B * slot, H, device="cuda", generator=g) * 2 + 0.5
        relb = torch.randn(H, 2 * R - 1, device="cuda", generator=g)
    outs = []
    for grouped in (0, 1):
        assert lib.ssr_tuning_set(b"attention_grouped", grouped) == 0
        out = torch.zeros(B * slot, D, device="cuda", dtype=torch.bfloat16)
        e = _err()
        rc = lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens_t.data_ptr(), _ptr(gate), _ptr(relb),
                               2 * R - 1, R - 1, 0, None, e, 512)
        torch.cuda.synchronize()
        lib.ssr_tuning_set(b"attention_grouped", 1)
        assert rc == 0, e.value.decode()
        outs.append(out)
    live = (torch.arange(slot, device="cuda")[None, :] < lens_t[:, None]).reshape(-1)
    assert torch.equal(outs[0][live], outs[1][live])
    ref = _attn_ref(qkv, B, slot, H, lens_t, gate, relb, R - 1)
    assert (outs[1].float() - ref)[live].abs().max().item() < 2e-2


@pytest.mark.parametrize("scale", [1.0, 6.0, 40.0])
def test_attention_stale_reference_and_rescale(scale):
    """Non-bias path (Whisper): O accumulates in TMEM against the block-0 reference maximum. Scores that keep rising
    along the key axis make that reference stale: mildly (probabilities far above 1, no rescale), strongly (partial
    sums leave the safe range -> in-place rescale of O and redo of the block), and some rows not at all."""
    lib = _lib()
    B, slot, H = 2, 700, 2
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(int(scale))
    qkv = torch.randn(B * slot, 3 * D, device="cuda", generator=g)
    q = qkv[:, :D].view(B, slot, H, 64)
    k = qkv[:, D:2 * D].view(B, slot, H, 64)
    q.mul_(0.05)
    # key j gets a component along e0 growing with j; queries have +1 / -1 / 0 along e0 depending on the row
    ramp = torch.linspace(0, 1, slot, device="cuda") * scale * 8
    k[..., 0] = ramp[None, :, None]
    sign = torch.tensor([1.0, -1.0, 0.0], device="cuda")[torch.arange(slot, device="cuda") % 3]
    q[..., 0] = sign[None, :, None]
    qkv = qkv.bfloat16()
    lens_t = torch.tensor([slot, 517], device="cuda", dtype=torch.int32)
    ref = _attn_ref(qkv, B, slot, H, lens_t, None, None, 0).view(B, slot, D)
    for variant in range(4):  # the rescale path under every arithmetic variant
        assert lib.ssr_tuning_set(b"attention_variant", variant) == 0
        out = torch.zeros(B * slot, D, device="cuda", dtype=torch.bfloat16)
        e = _err()
        rc = lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens_t.data_ptr(), None, None, 0, 0, 0,
                               None, e, 512)
        torch.cuda.synchronize()
        lib.ssr_tuning_set(b"attention_variant", DEFAULT_VARIANT)
        assert rc == 0, e.value.decode()
        got = out.float().view(B, slot, D)
        assert torch.isfinite(got).all()
        for b in range(B):
            L = int(lens_t[b])
            err = (got[b, :L] - ref[b, :L]).abs().max().item()
            assert err < 2e-2, f"variant {variant}, clip {b}: max err {err}"


def test_pool_mean():
    lib = _lib()
    B, slot, D = 5, 150, 1024
    x = torch.randn(B * slot, D, device="cuda")
    lens = torch.tensor([150, 149, 1, 77, 150], device="cuda", dtype=torch.int32)
    out = torch.zeros(B, 2, D, device="cuda")
    e = _err()
    rc = lib.ssr_pool_mean(x.data_ptr(), B, slot, D, lens.data_ptr(), out[:, 1].data_ptr(), 2 * D, None, e, 512)
    torch.cuda.synchronize()
    assert rc == 0, e.value.decode()
    for b in range(B):
        want = x.view(B, slot, D)[b, : int(lens[b])].mean(0)
        assert (out[b, 1] - want).abs().max().item() < 1e-5
