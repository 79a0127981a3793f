"""Level-0 small-K GEMM shapes of the UNet step, warm, for ncu --set full."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
x = rnd(32768, 320).bfloat16()
w320 = ops.pack_conv_weight(rnd(320, 320) / 18)
w960 = ops.pack_conv_weight(rnd(960, 320) / 18)
b320, b960 = rnd(320), rnd(960)
r32 = rnd(32768, 320)
x64 = rnd(8, 64, 64, 64).bfloat16()
wz = ops.pack_conv_weight(rnd(320, 64, 1, 1) / 8)
hf, hh = rnd(8, 64, 64, 320), rnd(8, 64, 64, 320).bfloat16()
fns = [
    lambda: ops.linear(x, w320, 320, bias=b320),                                   # plain bf16 out
    lambda: ops.linear(x, w320, 320, bias=b320, resid=r32, out_f32=True),          # to_out + fp32 residual
    lambda: ops.linear(x, w960, 960),                                              # qkv
    lambda: ops.conv_gemm(x64, wz, 320, 1, bias=b320, resid=hf, alpha=1.0, dual=True, out=(hf, hh)),  # zero-conv inject
]
for _ in range(3):
    for f in fns:
        f()
torch.cuda.synchronize()
for f in fns:   # warm timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record(); torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1)/20*1e3:.1f} us")
torch.cuda.cudart().cudaProfilerStart()
for f in fns:
    f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
