"""BASELINE.json configs 3 and 5 as smoke/throughput runs (synthetic inputs):
   C3 shape 768x512 (latent 64x96), C5 step sweep 2/5/10 at 512^2 batch 32."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import RDEIC, configs, synthetic  # noqa: E402
from rdeic_b200.pipeline import relay_decode  # noqa: E402

dev = torch.device("cuda:0")
params = configs.default_params()
model = RDEIC.from_config({"params": params}, device=dev)
model.load_state_dict(synthetic.make_state_dict(params, seed=231, device=dev))


def run(B, h, w, steps, sampler="ddpm", reps=2):
    g = torch.Generator(device=dev).manual_seed(0)
    cond = {"c_latent": [torch.randn(B, 4, h, w, generator=g, device=dev)],
            "c_crossattn": [torch.randn(B, 77, 1024, generator=g, device=dev)],
            "guide_hint": torch.randn(B, 256, h, w, generator=g, device=dev)}
    img = relay_decode(model, cond, steps, sampler=sampler)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        img = relay_decode(model, cond, steps, sampler=sampler)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    assert img.shape == (B, 8 * h, 8 * w, 3) and img.dtype == torch.uint8
    print(f"B={B} {8*h}x{8*w} steps={steps} {sampler}: {dt*1e3:.1f} ms/batch  {B/dt:.1f} images/s  "
          f"mem={torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)


run(16, 64, 96, 5)            # C3 shape (768 wide x 512 high), batch 16
run(64, 64, 96, 5, reps=1)    # C3 full batch 64 on one GPU
for s in (2, 5, 10):          # C5 sweep
    run(32, 64, 64, s, reps=1)
run(8, 64, 64, 5, sampler="ddim")
