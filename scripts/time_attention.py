"""Time the attention kernels on the shapes of a 512^2 batch-8 UNet + control step (warm, back to back).
Usage: python scripts/time_attention.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
B = 8
for name, heads, d, Nq, Nk in [("self L0", 5, 64, 4096, 4096), ("self L1", 10, 64, 1024, 1024), ("self L2", 20, 64, 256, 256),
                               ("self mid", 20, 64, 64, 64), ("cross L0", 5, 64, 4096, 77), ("cross L1", 10, 64, 1024, 77),
                               ("cross L2", 20, 64, 256, 77), ("ctrl self L0", 4, 16, 4096, 4096), ("ctrl self L1", 8, 16, 1024, 1024),
                               ("ctrl cross L0", 4, 16, 4096, 77), ("vae mid 512^2", 1, 512, 4096, 4096),
                               ("vae mid 768x512", 1, 512, 6144, 6144)]:
    C = heads * d
    qkv = torch.randn(B, Nq, 3 * C, generator=g, device=dev).bfloat16()
    kv = torch.randn(B, Nk, 2 * C, generator=g, device=dev).bfloat16()
    if Nq == Nk:
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q, k, v = qkv[..., :C], kv[..., :C], kv[..., C:]
    f = lambda: ops.attention(q, k, v, heads, d, d ** -0.5)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    fl = 4.0 * B * heads * Nq * Nk * d
    print(f"{name:14s} heads={heads:2d} d={d:2d} Nq={Nq:4d} Nk={Nk:4d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
