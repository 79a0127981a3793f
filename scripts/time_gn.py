import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops, _lib  # noqa: E402
import ctypes as C
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
lib = _lib.load()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for shape, dt in [((8, 64, 64, 320), torch.float32), ((8, 64, 64, 960), torch.float32), ((8, 32, 32, 1280), torch.float32),
                  ((8, 512, 512, 128), torch.bfloat16), ((8, 256, 256, 256), torch.bfloat16), ((8, 128, 128, 512), torch.bfloat16),
                  ((8, 64, 64, 320), torch.bfloat16)]:
    x = torch.randn(*shape, generator=g, device=dev).to(dt)
    C_ = shape[-1]
    ga, be = torch.ones(C_, device=dev), torch.zeros(C_, device=dev)
    us = t(lambda: ops.groupnorm(x, ga, be, 32, 1e-5, True))
    nbytes = x.numel() * x.element_size() + x.numel() * 2
    print(f"{shape} {str(dt)[6:]}: {us:8.1f} us  algorithmic {nbytes/us/1e3:7.0f} GB/s")
