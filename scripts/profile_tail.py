"""One launch of the fused VAE tail (norm_out + swish + conv_out + uint8) at 512^2, batch 8, for ncu."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
x128 = rnd(8, 512, 512, 128).bfloat16()
w11 = ops.pack_conv_weight(rnd(128, 128, 3, 3) / 34); b128 = rnd(128)
_, st = ops.conv_gemm(x128, w11, 128, 9, bias=b128, stats=True)
gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
wt = ops.pack_tail_weight(rnd(3, 128, 3, 3) / 34); b3 = rnd(3)
f = lambda: ops.gn_silu_conv3x3_tail(x128, st, gam, bet, 32, 1e-6, wt, b3, 3, True)
for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    f()
e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1)/10*1e3:.1f} us")
torch.cuda.cudart().cudaProfilerStart()
f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
