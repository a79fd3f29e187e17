"""Time the stage-level GEMM entry point at the encoder's shapes (CUDA events), for tuning and ncu captures."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import _lib

lib = _lib.load()
SHAPES = {  # name: (M, N, K, bias, act, resid, f32_out, bf16_out)
    "qkv": (38400, 3072, 1024, 1, 0, 0, 0, 1),
    "out": (38400, 1024, 1024, 1, 0, 1, 1, 0),
    "ffn1": (38400, 4096, 1024, 1, 1, 0, 0, 1),
    "ffn2": (38400, 1024, 4096, 1, 0, 1, 1, 0),
    "plain": (38400, 4096, 1024, 0, 0, 0, 0, 1),
    # out-proj variants: which part of the epilogue traffic costs what
    "out_f32_noresid": (38400, 1024, 1024, 1, 0, 0, 1, 0),
    "out_bf16_noresid": (38400, 1024, 1024, 1, 0, 0, 0, 1),
    "out_bf16_resid": (38400, 1024, 1024, 1, 0, 1, 0, 1),
}


def run(name, reps=5):
    M, N, K, hb, act, hr, f32, b16 = SHAPES[name]
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
    bias = torch.randn(N, device="cuda") if hb else None
    resid = torch.randn(M, N, device="cuda") if hr else None
    o32 = torch.empty(M, N, device="cuda") if f32 else None
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if b16 else None
    e = C.create_string_buffer(256)
    p = lambda t: None if t is None else t.data_ptr()
    st = torch.cuda.current_stream().cuda_stream

    def call():
        rc = lib.ssr_gemm_bf16(0, A.data_ptr(), K, M, W.data_ptr(), M, N, K, p(bias), act, p(resid), p(o32), p(o16), 0,
                               st, e, 256)
        assert rc == 0, e.value

    call()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        call()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    print(f"{name:6s} M={M} N={N} K={K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(SHAPES)
    for n in names:
        run(n)
