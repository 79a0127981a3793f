"""Run a handful of representative conv_gemm shapes (for `ncu --set full -k regex:conv_gemm`).
Each shape is launched twice (first = warm-up of lazy state) -> use -s/-c to pick launches."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402
from rdeic_b200.engine import Conv  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)


def geglu_linear():
    x = rnd(32768, 320).bfloat16()
    conv = Conv.load({"p.weight": rnd(2560, 320).cpu() / 18, "p.bias": rnd(2560).cpu()}, "p", dev, geglu=True)
    return lambda: ops.linear(x, conv.w, conv.n_out, bias=conv.b, act=2)


def proj_linear():
    x = rnd(32768, 320).bfloat16()
    w = ops.pack_conv_weight(rnd(320, 320) / 18)
    b, r = rnd(320), rnd(32768, 320)
    return lambda: ops.linear(x, w, 320, bias=b, resid=r, out_f32=True)


def vae_conv128():
    x = rnd(8, 512, 512, 128).bfloat16()
    w = ops.pack_conv_weight(rnd(128, 128, 3, 3) / 34)
    b, r = rnd(128), rnd(8, 512, 512, 128).bfloat16()
    return lambda: ops.conv_gemm(x, w, 128, 9, bias=b, resid=r)


def unet_conv320():
    x = rnd(8, 64, 64, 320).bfloat16()
    w = ops.pack_conv_weight(rnd(320, 320, 3, 3) / 54)
    b, r = rnd(320), rnd(8, 64, 64, 320)
    return lambda: ops.conv_gemm(x, w, 320, 9, bias=b, resid=r, dual=True)


fns = [geglu_linear(), proj_linear(), vae_conv128(), unet_conv320()]
for f in fns:
    f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for f in fns:
    f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
