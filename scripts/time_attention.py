"""Time the self-attention shapes of the UNet step (B=8): d=64, N in {4096, 1024, 256}."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for (B, heads, N) in [(8, 5, 4096), (8, 10, 1024), (8, 20, 256), (64, 5, 6144)][:3]:
    C = heads * 64
    qkv = torch.randn(B, N, 3 * C, device=dev).bfloat16()
    f = lambda: ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, 64, 0.125)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"B={B} heads={heads} N={N}: {us:.1f} us  {4.0 * B * heads * N * N * 64 / us / 1e6:.0f} TFLOP/s")
