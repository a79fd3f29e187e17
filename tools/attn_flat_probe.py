"""Whisper-shaped attention launch on near-flat scores (what a random-init model produces), for ncu captures."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssr_b200 import _lib

lib = _lib.load()
B, slot, H = 64, 1500, 20
D = H * 64
qkv = torch.randn(B * slot, 3 * D, device="cuda")
qkv[:, :D] *= 0.01
qkv = qkv.bfloat16()
lens = torch.full((B,), slot, device="cuda", dtype=torch.int32)
out = torch.zeros(B * slot, D, device="cuda", dtype=torch.bfloat16)
e = C.create_string_buffer(256)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    assert lib.ssr_attention(qkv.data_ptr(), out.data_ptr(), B, slot, H, lens.data_ptr(), None, None, 0, 0, 0, st, e, 256) == 0
torch.cuda.synchronize()
print("ok")
