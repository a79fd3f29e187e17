"""Per-item timeline of CTA 0 of the tcgen05 attention kernel (needs the -DSSR_ATT_TRACE build).

Prints, for the first items, SM-clock deltas between the barrier events of the three warp roles.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(HERE, "stuttering-speech-representation_b200", "build", "libssr_trace.so"))
lib.ssr_attention.restype = C.c_int32
lib.ssr_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_char_p, C.c_int32]
lib.ssr_att_trace_fetch.argtypes = [C.c_void_p]


def run(B, slot, H, length, bias):
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
    for _ in range(3):
        rc = lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens.data_ptr(), p(gate), p(rel), 2 * R - 1,
                               R - 1, 0, st, e, 256)
        assert rc == 0, e.value
        torch.cuda.synchronize()
    tr = np.zeros((3, 1024), np.int64)
    assert lib.ssr_att_trace_fetch(tr.ctypes.data_as(C.c_void_p)) == 0
    return tr


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "wavlm"
    if which == "wavlm":
        tr, nkb = run(256, 150, 16, 149, True), 3
    else:
        tr, nkb = run(8, 1500, 20, 1500, False), 24
    prod, mma, sm = tr
    t0 = sm[0]
    ps, ms, ss = 1 + nkb, 1 + 2 * nkb, 2 + 5 * nkb + 4  # events per item per role (mma: q_full, kv_full x nkb, bar_p x nkb)
    n_items = min(40, 1024 // ss)
    print("softmax warp 0 (cycles since first item start): loop top, start | per block: S ready, tmem ld done, exps done, proxy fence done, P arrived | O ready, O loaded, staged, epilogue done")
    for it in range(n_items):
        ev = sm[it * ss:(it + 1) * ss] - t0
        if ev[-1] <= 0 and it > 0:
            break
        print(f"item {it:2d}  " + " ".join(f"{int(v):7d}" for v in ev) + f"   total {int(ev[-1] - ev[0]):6d}")
    print("producer: q_empty passed | kv_empty passed x nkb")
    for it in range(min(n_items, 16)):
        ev = prod[it * ps:(it + 1) * ps] - t0
        print(f"item {it:2d}  " + " ".join(f"{int(v):7d}" for v in ev))
    print("mma thread: events in program order (q_full, then per block kv_full and — from the 2nd block on — bar_p of the previous block)")
    print(" ".join(f"{int(v - t0):7d}" for v in mma[:16 * ms]))


if __name__ == "__main__":
    main()
