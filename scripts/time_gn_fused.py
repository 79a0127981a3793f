"""Time conv_gemm with/without the fused GroupNorm statistics and groupnorm with/without them."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (B, H, W, C, f32) in [(8, 512, 512, 128, False), (8, 256, 256, 256, False), (8, 128, 128, 512, False), (8, 64, 64, 512, False),
                          (8, 64, 64, 320, True), (8, 32, 32, 640, True)]:
    x = torch.randn(B, H, W, C, device=dev).bfloat16()
    w = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / (3 * C ** 0.5))
    b = torch.randn(C, device=dev)
    r = torch.randn(B, H, W, C, device=dev)
    r = r if f32 else r.bfloat16()
    g, be = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    kw = dict(bias=b, resid=r, dual=f32)
    c0 = t(lambda: ops.conv_gemm(x, w, C, 9, **kw))
    c1 = t(lambda: ops.conv_gemm(x, w, C, 9, stats=True, **kw))
    out = ops.conv_gemm(x, w, C, 9, stats=True, **kw)
    y, st = out[0], out[-1]
    g0 = t(lambda: ops.groupnorm(y, g, be, 32, 1e-6, True))
    g1 = t(lambda: ops.groupnorm(y, g, be, 32, 1e-6, True, stats1=st))
    print(f"[{B},{H},{W},{C}] f32={f32}: conv {c0:.1f} -> {c1:.1f} us (+stats)   groupnorm {g0:.1f} -> {g1:.1f} us (from stats)")
