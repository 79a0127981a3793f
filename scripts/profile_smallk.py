"""Five launches for one ncu --set full capture: GEGLU K=320 N=2560, to_out K=320 N=320 + fp32 residual, then (0) qkv linear K=320 N=960 at M=32768 (plain bf16 out),
(1) the VAE 1x1 shortcut 256->128 at 512^2, (2) the VAE conv_out 128->4 (3x3) at 512^2."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
x = rnd(32768, 320).bfloat16(); w960 = ops.pack_conv_weight(rnd(960, 320) / 18)
x256 = rnd(8, 512, 512, 256).bfloat16(); wsc = ops.pack_conv_weight(rnd(128, 256, 1, 1) / 16); b128 = rnd(128)
x128 = rnd(8, 512, 512, 128).bfloat16(); wo = ops.pack_conv_weight(rnd(4, 128, 3, 3) / 34); b4 = rnd(4)
from rdeic_b200.engine import Conv  # noqa: E402
gg = Conv.load({"p.weight": rnd(2560, 320).cpu() / 18, "p.bias": rnd(2560).cpu()}, "p", dev, geglu=True)
w320 = ops.pack_conv_weight(rnd(320, 320) / 18); b320 = rnd(320); r32 = rnd(32768, 320)
fns = [lambda: ops.linear(x, gg.w, gg.n_out, bias=gg.b, act=2),
       lambda: ops.linear(x, w320, 320, bias=b320, resid=r32, out_f32=True),
       lambda: ops.linear(x, w960, 960),
       lambda: ops.conv_gemm(x256, wsc, 128, 1, bias=b128),
       lambda: ops.conv_gemm(x128, wo, 4, 9, bias=b4, out_f32=True)]
for _ in range(3):
    for f in fns:
        f()
torch.cuda.synchronize()
for f in fns:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record(); torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1)/10*1e3:.1f} us")
torch.cuda.cudart().cudaProfilerStart()
for f in fns:
    f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
