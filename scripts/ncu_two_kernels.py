"""Launch the level-0 self-attention and the level-0 conv3x3 of a 512^2 batch-8 step a few times (for `ncu --set full`)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, heads, d, N = 8, 5, 64, 4096
C = heads * d
qkv = torch.randn(B, N, 3 * C, generator=g, device=dev).bfloat16()
a = torch.randn(8, 64, 64, 320, generator=g, device=dev).bfloat16()
w = ops.pack_conv_weight((torch.randn(320, 320, 3, 3, generator=g, device=dev) / math.sqrt(2880)).contiguous())
resid = torch.randn(8, 64, 64, 320, generator=g, device=dev)
wl = ops.pack_conv_weight((torch.randn(2560, 320, generator=g, device=dev) / math.sqrt(320)).contiguous())
al = torch.randn(1, 1, 32768, 320, generator=g, device=dev).bfloat16()
for _ in range(3):
    ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, d, d ** -0.5)
    ops.conv_gemm(a, w, 320, 9, resid=resid, dual=True, stats=True)
    ops.conv_gemm(al, wl, 2560, 1, act=2)
torch.cuda.synchronize()
