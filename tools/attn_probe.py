"""Time the stage-level attention entry point (CUDA events) at the WavLM-Large / Whisper-large shapes."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import _lib

lib = _lib.load()
DEFAULT_VARIANT = 1  # csrc/attention_tc.cu g_attention_variant


def run(name, B, slot, H, length, bias, reps=5):
    D = H * 64
    qkv = torch.randn(B * slot, 3 * D, device="cuda").bfloat16()
    lens = torch.full((B,), length, device="cuda", dtype=torch.int32)
    R = 2048
    gate = torch.rand(B * slot, H, device="cuda") if bias else None
    rel = torch.randn(H, 2 * R - 1, device="cuda") if bias else None
    out = torch.zeros(B * slot, D, device="cuda", dtype=torch.bfloat16)
    e = C.create_string_buffer(256)
    p = lambda t: None if t is None else t.data_ptr()
    st = torch.cuda.current_stream().cuda_stream

    def call():
        rc = lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens.data_ptr(), p(gate), p(rel), 2 * R - 1,
                               R - 1, 0, st, e, 256)
        assert rc == 0, e.value

    call()
    torch.cuda.synchronize()
    # cross-check against the mma.sync kernel (impl = 1) at the full shape: exercises the persistent item walk
    ref = torch.zeros_like(out)
    rc = lib.ssr_attention(qkv.data_ptr(), ref.data_ptr(), B, slot, H, lens.data_ptr(), p(gate), p(rel), 2 * R - 1,
                           R - 1, 1, st, e, 256)
    assert rc == 0, e.value
    torch.cuda.synchronize()
    live = (torch.arange(slot, device="cuda")[None, :] < lens[:, None]).reshape(-1)
    diff = (out.float() - ref.float())[live].abs().max().item()
    print(f"{name:8s} max |tc - simt| over live rows = {diff:.3e}", flush=True)
    # (outputs are bf16: one ulp at |o| in [4, 8) is 3.1e-2; the GPU tests assert, this probe only reports)
    fl = 4.0 * B * H * length * length * 64
    combos = [(1, v) for v in (0, 1, 2, 3)]
    if 128 < slot <= 256:
        combos += [(0, v) for v in (0, 1)]
    combos += combos[:2]
    for paired, variant in combos:
        lib.ssr_tuning_set(b"attention_paired", paired)
        lib.ssr_tuning_set(b"attention_variant", variant)
        call()
        torch.cuda.synchronize()
        d = (out.float() - ref.float())[live].abs().max().item()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            call()
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        print(f"{name:8s} paired {paired} variant {variant} B={B} slot={slot} H={H}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s "
              f"(live)  max|tc - simt| {d:.3e}", flush=True)
    lib.ssr_tuning_set(b"attention_variant", DEFAULT_VARIANT)
    lib.ssr_tuning_set(b"attention_paired", 1)


if __name__ == "__main__":
    which = sys.argv[1:] or ["wavlm", "whisper"]
    if "wavlm" in which:
        run("wavlm", 256, 150, 16, 149, True, reps=20)
    if "whisper" in which:
        run("whisper", 64, 1500, 20, 1500, False, reps=6)
