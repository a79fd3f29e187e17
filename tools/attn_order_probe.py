"""A/B of the attention item orders (knob "attention_grouped") on B200: the outputs must be bit-identical (the work
per item is the same, only who does it when changes), kernel alone at the Whisper-large shape, then the whole
Whisper-large step. Writes lines as it goes (short GPU slots)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import WhisperEncoderEngine, _lib, synth

lib = _lib.load()
T0 = time.time()


def say(*a):
    print(f"[{time.time() - T0:6.1f}s]", *a, flush=True)


def attention(B, slot, H, lens, bias, reps):
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + slot)
    qkv = torch.randn(B * slot, 3 * D, device="cuda", generator=g).bfloat16()
    lens_d = torch.tensor(lens, device="cuda", dtype=torch.int32)
    R = 2048
    gate = torch.rand(B * slot, H, device="cuda", generator=g) if bias else None
    rel = torch.randn(H, 2 * R - 1, device="cuda", generator=g) if bias else None
    e = C.create_string_buffer(256)
    p = lambda t: None if t is None else t.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    outs, times = {}, {0: [], 1: []}
    live = (torch.arange(slot, device="cuda")[None, :] < lens_d[:, None]).reshape(-1)
    for grouped in (0, 1, 0, 1):
        assert lib.ssr_tuning_set(b"attention_grouped", grouped) == 0
        out = torch.zeros(B * slot, D, device="cuda", dtype=torch.bfloat16)

        def call():
            rc = lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens_d.data_ptr(), p(gate), p(rel),
                                   2 * R - 1, R - 1, 0, st, e, 256)
            assert rc == 0, e.value

        call()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            call()
        ev[1].record()
        torch.cuda.synchronize()
        times[grouped].append(ev[0].elapsed_time(ev[1]) / reps * 1e3)
        outs.setdefault(grouped, out)
        assert torch.equal(out[live], outs[grouped][live])
    same = torch.equal(outs[0][live], outs[1][live])
    say(f"B={B} slot={slot} H={H} bias={bias} lens={lens[:4]}..: bit-identical {same}; us major {times[0]} grouped {times[1]}")
    lib.ssr_tuning_set(b"attention_grouped", 1)
    return same


ok = True
for args in [(64, 1500, 20, [1500] * 64, False, 6), (5, 1500, 20, [1500, 777, 129, 1, 1290], False, 2),
             (7, 385, 3, [385, 384, 257, 256, 129, 128, 5], True, 2),
             (33, 640, 12, [640 - 19 * i for i in range(33)], True, 2), (40, 640, 4, [640 - 15 * i for i in range(40)], False, 2)]:
    try:
        ok &= attention(*args)
    except Exception as ex:  # keep going: the step A/B below is the number that matters
        ok = False
        say("FAILED", args[:3], repr(ex))
say("ALL BIT-IDENTICAL" if ok else "MISMATCH / FAILURE")

enc, wfe = synth.build_whisper_encoder("large", seed=0)
eng = WhisperEncoderEngine.from_hf(enc, wfe, device=0)
say("engine built")
WB = 64
audio = torch.from_numpy(np.stack([synth.clip_by_index(i, 48000) for i in range(WB)])).cuda()
n = np.full(WB, 48000, np.int32)
res = {}
for grouped in (0, 1, 0, 1):
    lib.ssr_tuning_set(b"attention_grouped", grouped)
    out = eng.pooled_device(audio, n)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(4):
        out = eng.pooled_device(audio, n, out=out)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 4
    say(f"whisper-large step, grouped {grouped}: {ms:.2f} ms/step = {WB / ms * 1e3:.1f} clips/s")
    if grouped in res:
        assert torch.equal(res[grouped], out)
    res[grouped] = out.clone()
say("step outputs bit-identical:", torch.equal(res[0], res[1]))
