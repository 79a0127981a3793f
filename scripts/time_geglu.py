"""GEGLU projections (K -> 8K, GELU gate in the epilogue) and small-K linears of UNet levels 0-2 at batch 8:
tile width / epilogue warp count A/B.  RDEIC_TILE_OVERHEAD sets the tile-picker constant."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402
from rdeic_b200.engine import Conv  # noqa: E402
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for M, C in [(32768, 320), (8192, 640), (2048, 1280)]:
    x = rnd(M, C).bfloat16()
    gg = Conv.load({"p.weight": rnd(8 * C, C).cpu() / C ** 0.5, "p.bias": rnd(8 * C).cpu()}, "p", dev, geglu=True)
    fl = 2.0 * M * C * 8 * C
    for tn in (0, 160, 256):
        us = t(lambda: ops.linear(x, gg.w, gg.n_out, bias=gg.b, act=2, tile_n=tn))
        print(f"geglu M={M} K={C} tile_n={tn:3d} {us:7.1f} us {fl/us/1e6:7.1f} TF/s", flush=True)
    w = ops.pack_conv_weight(rnd(C, C) / C ** 0.5); b = rnd(C); r32 = rnd(M, C)
    w3 = ops.pack_conv_weight(rnd(3 * C, C) / C ** 0.5)
    for name, fn, n in [("to_out+resid f32", lambda: ops.linear(x, w, C, bias=b, resid=r32, out_f32=True), C),
                        ("qkv", lambda: ops.linear(x, w3, 3 * C), 3 * C)]:
        for tn in (0, 160, 256):
            us = t(lambda: fn() if tn == 0 else (ops.linear(x, w, C, bias=b, resid=r32, out_f32=True, tile_n=tn) if n == C else ops.linear(x, w3, 3 * C, tile_n=tn)))
            print(f"{name:18s} M={M} K={C} tile_n={tn:3d} {us:7.1f} us {2.0*M*C*n/us/1e6:7.1f} TF/s", flush=True)
