"""A/B the VAE decode and the whole relay decode with and without the CUDA-graph replay of the VAE
(RDEIC.VAE_GRAPH_MAX_POSITIONS), and check the two produce identical images.
Usage: python scripts/ab_vae_graph.py [batch]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import H, W, make_inputs  # noqa: E402
from rdeic_b200 import RDEIC, configs, synthetic  # noqa: E402
from rdeic_b200.pipeline import relay_decode  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
params = configs.default_params()
model = RDEIC.from_config({"params": params}, device=dev)
model.load_state_dict(synthetic.make_state_dict(params, seed=231, device=dev))
c_latent, hint, ctx, noises = make_inputs(batch, H // 8, W // 8)
d = lambda t: t.to(dev)
cond = {"c_latent": [d(c_latent)], "c_crossattn": [d(ctx)], "guide_hint": d(hint)}
z = d(c_latent)
ns = [d(n) for n in noises]


def timeit(fn, n):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


full = lambda: relay_decode(model, cond, 5, start_noise=ns[0], step_noises=ns[1:])
res = {}
for limit in (1 << 30, 0, 1 << 30, 0):
    model.VAE_GRAPH_MAX_POSITIONS = limit
    img = model.decode_first_stage_u8(z).clone()
    res.setdefault(limit > 0, img)
    print(f"vae graph {'on ' if limit else 'off'}: vae_ms={timeit(lambda: model.decode_first_stage_u8(z), 9):.3f} "
          f"decode_ms={timeit(full, 7):.3f}", flush=True)
print("identical images:", bool(torch.equal(res[True], res[False])))
