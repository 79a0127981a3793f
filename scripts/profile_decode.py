"""One eager (un-graphed) 512x512 batch-8 relay decode between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Usage: python scripts/profile_decode.py [steps] [batch] [part]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import BATCH, H, W, make_inputs  # noqa: E402
from rdeic_b200 import RDEIC, configs, synthetic  # noqa: E402
from rdeic_b200.pipeline import relay_decode  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batch = int(sys.argv[2]) if len(sys.argv) > 2 else BATCH
part = sys.argv[3] if len(sys.argv) > 3 else "all"
dev = torch.device("cuda:0")
params = configs.default_params()
model = RDEIC.from_config({"params": params}, device=dev, use_cuda_graph=False)
model.load_state_dict(synthetic.make_state_dict(params, seed=231, device=dev))
c_latent, hint, ctx, noises = make_inputs(batch, H // 8, W // 8)
d = lambda t: t.to(dev)
cond = {"c_latent": [d(c_latent)], "c_crossattn": [d(ctx)], "guide_hint": d(hint)}
ns = [d(n) for n in noises]
tt = torch.full((batch,), 224, dtype=torch.long, device=dev)


def run():
    if part == "unet":
        return model.apply_model(ns[0], tt, cond)
    if part == "vae":
        return model.decode_first_stage_u8(d(c_latent))
    return relay_decode(model, cond, steps, start_noise=ns[0], step_noises=ns[1:])


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
