"""Per-kernel device time of one CUDA-graph UNet+control step and one VAE decode, from CUPTI through torch.profiler
(no kernel replay, so the graph's two-stream overlap and PDL stay as in production): where the step goes now.
A breakdown tool; bench numbers never come from a profiled run.  Usage: python scripts/prof_step.py [batch]"""
import re
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

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
x, z = d(noises[0]), d(c_latent)
for _ in range(3):
    model.apply_model(x, tt, cond)
    model.decode_first_stage_u8(z)
torch.cuda.synchronize()


def run(fn, n, title):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    t0, t1 = 1e30, 0.0
    for e in prof.events():
        if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.device_time_total > 0:
            name = re.sub(r"^void ", "", e.name).replace("rdeic::", "")
            name = re.sub(r"\(.*", "", name)
            agg[name][0] += e.device_time_total
            agg[name][1] += 1
            t0, t1 = min(t0, e.time_range.start), max(t1, e.time_range.end)
    tot = sum(v[0] for v in agg.values())
    print(f"== {title}: {n} x; sum of kernel time {tot / n / 1e3:.3f} ms per call, span {(t1 - t0) / n / 1e3:.3f} ms per call")
    for name, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"  {us / n / 1e3:8.3f} ms {100 * us / tot:5.1f}%  n={cnt // n:4d}  avg {us / cnt:8.1f} us  {name[:110]}")


run(lambda: model.apply_model(x, tt, cond), 5, f"UNet+control step, batch {batch} (graph replay)")
run(lambda: model.decode_first_stage_u8(z), 3, f"VAE decode to uint8, batch {batch}")
