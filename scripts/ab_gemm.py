"""A/B a handful of plain conv_gemm shapes of the UNet step for the library selected by RDEIC_B200_LIB (two builds of
csrc/ timed on one box: boxes differ by a few per cent, builds must not).  Usage: python scripts/ab_gemm.py"""
import math
import os
import sys
from pathlib import Path

import torch

os.environ.setdefault("RDEIC_B200_LIB_PARTIAL", "1")       # an older build lacks the newest symbols
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def bench(name, a_shape, n_out, taps, **kw):
    N, H, W, C = a_shape
    a = torch.randn(a_shape, generator=g, device=dev).bfloat16()
    k = int(math.isqrt(taps))
    w = ops.pack_conv_weight((torch.randn(n_out, C, k, k, generator=g, device=dev) / math.sqrt(taps * C)).contiguous())
    bias = torch.randn(n_out, generator=g, device=dev)
    if kw.pop("resid32", False):
        kw["resid"] = torch.randn(N, H, W, n_out, generator=g, device=dev)
    f = lambda: ops.conv_gemm(a, w, n_out, taps, bias=bias, **kw)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    print(f"{_lib.LIB_PATH.parent.name}/{_lib.LIB_PATH.name:24s} {name:34s} {e0.elapsed_time(e1) * 50:8.1f} us")


bench("GEGLU L0 K=320 N=2560", (1, 1, 32768, 320), 2560, 1, act=2)
bench("qkv L0 K=320 N=960", (1, 1, 32768, 320), 960, 1)
bench("to_out L0 f32 resid", (1, 1, 32768, 320), 320, 1, resid32=True, out_f32=True)
bench("proj_in L0 f32", (1, 1, 32768, 320), 320, 1, out_f32=True)
bench("ff2 L0 K=1280 resid", (1, 1, 32768, 1280), 320, 1, resid32=True)
bench("conv3x3 L0 320 resid dual stats", (8, 64, 64, 320), 320, 9, resid32=True, dual=True, stats=True)
bench("conv3x3 L0 320 stats", (8, 64, 64, 320), 320, 9, stats=True)
bench("conv3x3 L1 640 resid dual stats", (8, 32, 32, 640), 640, 9, resid32=True, dual=True, stats=True)
bench("conv3x3 L2 1280 resid dual stats", (8, 16, 16, 1280), 1280, 9, resid32=True, dual=True, stats=True)
bench("conv3x3 L3 1280 (split-K)", (8, 8, 8, 1280), 1280, 9, resid32=True, dual=True, stats=True)
bench("GEGLU L1 K=640 N=5120", (1, 1, 8192, 640), 5120, 1, act=2)
bench("linear L1 f32 resid", (1, 1, 8192, 640), 640, 1, resid32=True, out_f32=True)
bench("VAE conv 512 @128^2", (8, 128, 128, 512), 512, 9, stats=True)
bench("VAE conv 128 @512^2 resid", (8, 512, 512, 128), 128, 9, stats=True)
