import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, H, W, C) in [(8, 512, 512, 128), (8, 256, 256, 256)]:
    x = rnd(B, H, W, C).bfloat16()
    w = ops.pack_conv_weight(rnd(C, C, 3, 3) / (3 * C ** 0.5))
    b = rnd(C)
    r16 = rnd(B, H, W, C).bfloat16()
    r32 = r16.float()
    fl = 2.0 * B * H * W * C * 9 * C
    for name, kw in [("no resid", {}), ("bf16 resid", dict(resid=r16)), ("f32 resid", dict(resid=r32)),
                     ("f32 resid dual", dict(resid=r32, dual=True))]:
        us = t(lambda: ops.conv_gemm(x, w, C, 9, bias=b, **kw))
        print(f"[{B},{H},{W},{C}] {name:16s} {us:8.1f} us  {fl/us/1e6:7.1f} TF/s", flush=True)
# N = 128 layers of the 512^2 level (two M tiles per item share the weight stage; RDEIC_DUAL_M=0 disables)
x256 = rnd(8, 512, 512, 256).bfloat16()
x128 = rnd(8, 512, 512, 128).bfloat16()
w21 = ops.pack_conv_weight(rnd(128, 256, 3, 3) / 48)
w11 = ops.pack_conv_weight(rnd(128, 128, 3, 3) / 34)
wo = ops.pack_conv_weight(rnd(4, 128, 3, 3) / 34)
b128, b4 = rnd(128), rnd(4)
for name, fn, fl in [("256->128 stats", lambda: ops.conv_gemm(x256, w21, 128, 9, bias=b128, stats=True), 2.0 * 8 * 512 * 512 * 128 * 9 * 256),
                     ("128->128 stats+resid", lambda: ops.conv_gemm(x128, w11, 128, 9, bias=b128, resid=x128, stats=True), 2.0 * 8 * 512 * 512 * 128 * 9 * 128),
                     ("128->4 conv_out f32", lambda: ops.conv_gemm(x128, wo, 4, 9, bias=b4, out_f32=True), 2.0 * 8 * 512 * 512 * 4 * 9 * 128)]:
    us = t(fn)
    print(f"{name:24s} {us:8.1f} us  {fl/us/1e6:7.1f} TF/s", flush=True)
# fused tail: norm_out + swish + conv_out + uint8 vs the three-kernel path
_, st = ops.conv_gemm(x128, w11, 128, 9, bias=b128, stats=True)
gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
wt = ops.pack_tail_weight(rnd(3, 128, 3, 3) / 34); b3 = rnd(3)
def unfused():
    xn = ops.groupnorm(x128, gam, bet, 32, 1e-6, True, stats1=st)
    return ops.image_to_u8(ops.conv_gemm(xn, wo, 4, 9, bias=b4, out_f32=True))
print(f"tail unfused (gn fold+apply, conv_out, u8) {t(unfused):8.1f} us", flush=True)
print(f"tail fused                                 {t(lambda: ops.gn_silu_conv3x3_tail(x128, st, gam, bet, 32, 1e-6, wt, b3, 3, True)):8.1f} us", flush=True)
