import ctypes as C, os, sys, torch
sys.path.insert(0, "/root/repo")
from ssr_b200 import _lib
lib = _lib.load()
B, slot, H = 64, 1500, 20
D = H * 64
for qs in (1.0, 0.125, 0.01, 0.0):
    qkv = torch.randn(B * slot, 3 * D, device="cuda")
    qkv[:, :D] *= qs
    qkv = qkv.bfloat16()
    lens = torch.full((B,), slot, device="cuda", dtype=torch.int32)
    out = torch.zeros(B * slot, D, device="cuda", dtype=torch.bfloat16)
    e = C.create_string_buffer(256)
    st = torch.cuda.current_stream().cuda_stream
    def call():
        assert lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens.data_ptr(), None, None, 0, 0, 0, st, e, 256) == 0
    call(); torch.cuda.synchronize()
    for reps in (2, 20):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps): call()
        ev[1].record(); torch.cuda.synchronize()
        print(f"q scale {qs}: reps {reps}: {ev[0].elapsed_time(ev[1]) / reps * 1e3:.1f} us", flush=True)
