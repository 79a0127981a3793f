"""bf16 throughput mode vs fp32 kernel mode (1e-6 of the reference) for several weight / input seeds:
per-step UNet rel-L2 at full width, latent 32x32 and 64x64.  Prints one line per seed."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import rel_l2  # noqa: E402
from rdeic_b200 import RDEIC, configs, synthetic  # noqa: E402

dev = torch.device("cuda:0")
params = configs.default_params()
for seed in [int(a) for a in sys.argv[1:]] or [231, 1, 2, 3]:
    sd = synthetic.make_state_dict(params, seed=seed, device=dev)
    m16 = RDEIC.from_config({"params": params}, device=dev, use_cuda_graph=False).load_state_dict(sd)
    m32 = RDEIC.from_config({"params": params}, device=dev, precision="fp32").load_state_dict(sd)
    out = []
    for (B, h) in ((1, 32), (2, 64)):
        g = torch.Generator(device=dev).manual_seed(seed + 100)
        x = torch.randn(B, 4, h, h, generator=g, device=dev)
        cond = {"c_latent": [x], "c_crossattn": [torch.randn(B, 77, 1024, generator=g, device=dev)],
                "guide_hint": torch.randn(B, 256, h, h, generator=g, device=dev)}
        t = torch.full((B,), 224, dtype=torch.long, device=dev)
        e = rel_l2(m16.apply_model(x, t, cond).cpu().numpy(), m32.apply_model(x, t, cond).cpu().numpy())
        eu = rel_l2(m16.apply_model_unconditional(x, t, cond).cpu().numpy(),
                    m32.apply_model_unconditional(x, t, cond).cpu().numpy())
        out.append(f"{h * 8}^2: cond {e:.3e} uncond {eu:.3e}")
    print(f"seed {seed}: " + " | ".join(out), flush=True)
    del m16, m32, sd
    torch.cuda.empty_cache()
