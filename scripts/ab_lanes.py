"""UNet+control step as one batch or as two half-batch lanes on two streams: same result?  which is faster?
Usage: python scripts/ab_lanes.py [batch]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import H, W, make_inputs  # noqa: E402
from rdeic_b200 import RDEIC, configs, synthetic  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
params = configs.default_params()
model = RDEIC.from_config({"params": params}, device=dev)
model.load_state_dict(synthetic.make_state_dict(params, seed=231, device=dev))
c_latent, hint, ctx, noises = make_inputs(batch, H // 8, W // 8)
d = lambda t: t.to(dev)
cond = {"c_latent": [d(c_latent)], "c_crossattn": [d(ctx)], "guide_hint": d(hint)}
tt = torch.full((batch,), 224, dtype=torch.long, device=dev)
x = d(noises[0])


def timeit(fn, n):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


outs = {}
for lanes in (1, 2, 1, 2):
    model.control_model.lanes = lanes
    model._graphs.clear()
    outs[lanes] = model.apply_model(x, tt, cond).float()
    ms = timeit(lambda: model.apply_model(x, tt, cond), 15)
    print(f"lanes={lanes}: unet_step_ms={ms:.3f}", flush=True)
rel = ((outs[1] - outs[2]).norm() / outs[1].norm()).item()
print(f"lanes 2 vs 1: rel-L2 {rel:.3e}")
