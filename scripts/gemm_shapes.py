"""Collect every conv_gemm call of one eager decode (shape + epilogue), then time each unique
shape in isolation (CUDA events, L2 flushed between launches) and print TFLOP/s per shape.
Usage: python scripts/gemm_shapes.py [unet|vae] [batch] [out.json]"""
import json
import sys
from collections import OrderedDict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import H, W, make_inputs  # noqa: E402
from rdeic_b200 import RDEIC, configs, ops, synthetic  # noqa: E402

part = sys.argv[1] if len(sys.argv) > 1 else "unet"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
out_path = sys.argv[3] if len(sys.argv) > 3 else None
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
import rdeic_b200.engine as eng  # noqa: E402
if part == "unet":
    model.apply_model(d(noises[0]), tt, cond)
else:
    model.decode_first_stage_u8(d(c_latent))
torch.cuda.synchronize()
ops.conv_gemm = orig

uniq = OrderedDict()
for a, w, n_out, taps, kw in calls:
    a2 = kw.get("a2")
    key = (tuple(a.shape), None if a2 is None else a2.shape[-1], n_out, taps, kw.get("resid") is not None,
           bool(kw.get("dual")), bool(kw.get("out_f32")), kw.get("w_batch_stride", 0) != 0)
    uniq.setdefault(key, []).append((a, w, n_out, taps, kw))

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for key, lst in uniq.items():
    a, w, n_out, taps, kw = lst[0]
    kw = dict(kw)
    kw.pop("out", None)
    # warm, back-to-back (what the kernel sees inside the CUDA-graphed step: producers leave their
    # outputs in the 126 MB L2); a flushed measurement mostly times the write-back of the flush buffer
    for _ in range(2):
        orig(a, w, n_out, taps, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        orig(a, w, n_out, taps, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    N, Hh, Ww, C = a.shape
    k_true = taps * (C + (key[1] or 0))
    fl = 2.0 * N * Hh * Ww * n_out * k_true
    rows.append(dict(a=list(a.shape), a2=key[1], n_out=n_out, taps=taps, resid=key[4], dual=key[5], f32=key[6],
                     count=len(lst), us=ms * 1e3, tflops=fl / ms / 1e9, gflop=fl / 1e9))
tot = sum(r["us"] * r["count"] for r in rows)
print(f"{part} B={batch}: {len(calls)} calls, {len(rows)} unique shapes, sum(warm, isolated) = {tot/1e3:.3f} ms")
for r in sorted(rows, key=lambda r: -r["us"] * r["count"]):
    print(f"{r['us']*r['count']/1e3:7.3f} ms n={r['count']:3d} {r['us']:8.1f} us {r['tflops']:7.1f} TF/s  M={r['a'][0]*r['a'][1]*r['a'][2]:7d} "
          f"A={r['a']} a2={r['a2']} N={r['n_out']} taps={r['taps']} resid={int(r['resid'])} dual={int(r['dual'])} f32={int(r['f32'])}")
if out_path:
    Path(out_path).write_text(json.dumps(rows))
