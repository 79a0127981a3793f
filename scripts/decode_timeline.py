"""Where does one decode (batch 8, 512^2, 5 relay steps) spend its time?  CUDA events at the segment
boundaries of back-to-back decodes (no host sync inside the loop): q_sample | step 0..4 | VAE."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import BATCH, H, W, RELAY_STEPS, WEIGHT_SEED, make_inputs  # noqa: E402
from rdeic_b200 import RDEIC, configs, synthetic  # noqa: E402
from rdeic_b200.pipeline import relay_decode  # noqa: E402
from rdeic_b200.spaced_sampler_relay import SpacedSampler  # noqa: E402

dev = torch.device("cuda:0")
params = configs.default_params()
model = RDEIC.from_config({"params": params}, device=dev)
model.load_state_dict(synthetic.make_state_dict(params, seed=WEIGHT_SEED, device=dev))
c_latent, hint, ctx, noises = make_inputs(BATCH, H // 8, W // 8)
d = lambda t: t.to(dev)
cond = {"c_latent": [d(c_latent)], "c_crossattn": [d(ctx)], "guide_hint": d(hint)}
nz = [d(n) for n in noises]
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
orig_step = SpacedSampler.p_sample_spaced
def step(self, *a, **k):
    out = orig_step(self, *a, **k); mark(f"step{k.get('step_i', 0)}"); return out
SpacedSampler.p_sample_spaced = step
orig_vae = model.decode_first_stage_u8
def vae(z):
    mark("pre_vae"); out = orig_vae(z); mark("vae"); return out
model.decode_first_stage_u8 = vae
run = lambda: relay_decode(model, cond, RELAY_STEPS, sampler="ddpm", start_noise=nz[0], step_noises=nz[1:])
for _ in range(3):
    run()
torch.cuda.synchronize()
marks.clear()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
mark("start")
for _ in range(n):
    run(); 
torch.cuda.synchronize()
acc = {}
for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
    acc.setdefault(n1 if n1 != "step0" else "q_sample+step0", []).append(e0.elapsed_time(e1))
tot = marks[0][1].elapsed_time(marks[-1][1]) / n
for k, v in acc.items():
    print(f"{k:16s} mean {sum(v)/len(v):7.3f} ms  min {min(v):7.3f}  max {max(v):7.3f}")
print(f"total per decode {tot:.3f} ms")
