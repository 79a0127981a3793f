"""What does a narrow tile cost?  One Linear, M = 148 * 128 rows (one M tile per SM), K = 2048 (32 k-blocks per
tile, A = 77 MB stays in L2 across repetitions), N = 2048, run with every tile width the kernel has: every CTA
walks 2048 / tile_n tiles of 128 MMAs each.  If an M = 128 tcgen05.mma cost N/2 clocks the time would be flat in
the width (same FLOPs); the measured times say how far it is from that."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
M, K, N = 148 * 128, 2048, 2048
x = (torch.randn(M, K, generator=g, device=dev) / 8).bfloat16()
w = ops.pack_conv_weight(torch.randn(N, K, generator=g, device=dev) / 64)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
fl = 2.0 * M * K * N
for tn in (256, 160, 128, 64, 32):
    us = t(lambda: ops.linear(x, w, N, tile_n=tn))
    per_cta = -(-N // tn) * (K // 64) * 4
    print(f"tile_n={tn:3d}: {us:7.1f} us  {fl/us/1e6:7.1f} TF/s  {per_cta} MMAs per CTA -> {us*1e-6*1.9e9/per_cta:.0f} clk per MMA at 1.9 GHz "
          f"(incl. launch + pipeline fill)", flush=True)
