"""Time every distinct conv_gemm shape of one UNet+control step (or VAE decode) at each tile width the kernel has
(tile_n hint 128 / 160 / 256) next to the library's own choice, warm and back to back.  Input for `pick_block_n`.
Usage: python scripts/tile_sweep.py [unet|vae] [batch]"""
import sys
from collections import OrderedDict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import H, W, make_inputs  # noqa: E402
from rdeic_b200 import RDEIC, configs, ops, synthetic  # noqa: E402

part = sys.argv[1] if len(sys.argv) > 1 else "unet"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
params = configs.default_params()
model = RDEIC.from_config({"params": params}, device=dev, use_cuda_graph=False)
model.load_state_dict(synthetic.make_state_dict(params, seed=231, device=dev))
c_latent, hint, ctx, noises = make_inputs(batch, H // 8, W // 8)
d = lambda t: t.to(dev)
cond = {"c_latent": [d(c_latent)], "c_crossattn": [d(ctx)], "guide_hint": d(hint)}
tt = torch.full((batch,), 224, dtype=torch.long, device=dev)
calls = []
orig = ops.conv_gemm


def spy(a, w_packed, n_out, taps, **kw):
    calls.append((a, w_packed, n_out, taps, dict(kw)))
    return orig(a, w_packed, n_out, taps, **kw)


ops.conv_gemm = spy
if part == "unet":
    model.apply_model(d(noises[0]), tt, cond)
else:
    model.decode_first_stage_u8(d(c_latent))
torch.cuda.synchronize()
ops.conv_gemm = orig
uniq = OrderedDict()
for a, w, n_out, taps, kw in calls:
    a2 = kw.get("a2")
    key = (tuple(a.shape), None if a2 is None else a2.shape[-1], n_out, taps, kw.get("resid") is not None, bool(kw.get("dual")),
           bool(kw.get("out_f32")), kw.get("act", 0), bool(kw.get("up2")), bool(kw.get("stride2")), kw.get("w2") is not None,
           bool(kw.get("stats")))
    uniq.setdefault(key, []).append((a, w, n_out, taps, kw))


def time_it(a, w, n_out, taps, kw, tile):
    kw = dict(kw)
    kw.pop("out", None)
    kw["tile_n"] = tile
    for _ in range(2):
        orig(a, w, n_out, taps, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        orig(a, w, n_out, taps, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 100            # us per launch


rows = []
for key, lst in uniq.items():
    a, w, n_out, taps, kw = lst[0]
    t = {tile: time_it(a, w, n_out, taps, kw, tile) for tile in (0, 128, 160, 256)}
    rows.append((len(lst), key, t))
tot = {tile: sum(n * t[tile] for n, _, t in rows) for tile in (0, 128, 160, 256)}
best = sum(n * min(t.values()) for n, _, t in rows)
print(f"{part} B={batch}: {len(calls)} calls, {len(rows)} shapes; sum us: auto {tot[0]:.0f}  128 {tot[128]:.0f}  160 {tot[160]:.0f}  "
      f"256 {tot[256]:.0f}  best-of {best:.0f}")
for n, key, t in sorted(rows, key=lambda r: -(r[2][0] - min(r[2].values())) * r[0]):
    shape, a2, n_out, taps, resid, dual, f32, act, up2, s2, inj, st = key
    gain = (t[0] - min(t.values())) * n
    print(f"gain {gain:7.1f} us n={n:3d} auto {t[0]:7.1f} | 128 {t[128]:7.1f} | 160 {t[160]:7.1f} | 256 {t[256]:7.1f}  A={list(shape)} a2={a2} "
          f"N={n_out} taps={taps} resid={int(resid)} dual={int(dual)} f32={int(f32)} act={act} up2={int(up2)} s2={int(s2)} inj={int(inj)} st={int(st)}")
