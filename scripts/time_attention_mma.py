"""Time the attention shapes that run on the mma.sync kernel in a UNet+control step (B=8): the control
adapter's d = 16 self-attention (N = 4096 / 1024 / 256, heads 4 / 8 / 16) and the d = 64 / d = 16
cross-attention over the 77 text tokens.  RDEIC_B200_LIB selects the build (A/B on one box)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
out = []
for (B, heads, d, Nq, Nk) in [(8, 4, 16, 4096, 4096), (8, 8, 16, 1024, 1024), (8, 16, 16, 256, 256),
                               (8, 5, 64, 4096, 77), (8, 4, 16, 4096, 77)]:
    C = heads * d
    q = torch.randn(B, Nq, C, device=dev).bfloat16()
    kv = torch.randn(B, Nk, 2 * C, device=dev).bfloat16()
    f = lambda: ops.attention(q, kv[..., :C], kv[..., C:], heads, d, d ** -0.5)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    out.append(f"d={d} h={heads} {Nq}x{Nk}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
print(_lib.LIB_PATH.name, " | ".join(out))
