"""Where one CUDA-graph UNet+control step goes, kernel by kernel, from CUPTI through torch.profiler (no kernel replay, so
the graph's two-stream overlap and PDL stay as in production).  With programmatic dependent launch a kernel's recorded
duration includes the time it sits at griddepcontrol.wait behind its predecessor, so durations are not additive; what is
additive along a stream is each kernel's END-TO-END increment end(i) - end(i-1).  The script prints, per kernel class on
the stream that carries the critical path, the sum of those increments (= that class's share of the step).
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


def short(name):
    name = re.sub(r"^void ", "", name).replace("rdeic::", "")
    return re.sub(r"\(.*", "", name)


def run(fn, title):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.profiler.kineto_results.events() if str(e.device_type()).lower().endswith("cuda") and e.duration_ns() > 0]
    streams = defaultdict(list)
    for e in evs:
        streams[e.device_resource_id()].append((e.start_ns(), e.start_ns() + e.duration_ns(), short(e.name())))
    t0 = min(s for v in streams.values() for s, _, _ in v)
    t1 = max(t for v in streams.values() for _, t, _ in v)
    print(f"== {title}: span {(t1 - t0) / 1e6:.3f} ms, {len(evs)} kernels on {len(streams)} stream(s)")
    # A CUDA graph spreads its branches over internal streams, so per-stream gaps mean nothing.  Sweep the union
    # timeline instead: time with 0 / 1 / >= 2 kernels in flight, and for every instant with exactly one kernel in
    # flight charge it to that kernel (the serial part of the step); overlapped time is charged to "(overlap)".
    pts = []
    allk = [k for v in streams.values() for k in v]
    for i, (s_, t_, name) in enumerate(allk):
        pts.append((s_, 1, i))
        pts.append((t_, -1, i))
    pts.sort()
    live = set()
    last = pts[0][0]
    excl = defaultdict(float)
    hist = defaultdict(float)
    for t_, kind, i in pts:
        dt = t_ - last
        if dt > 0:
            n = len(live)
            hist[min(n, 3)] += dt
            if n == 0:
                excl["(no kernel in flight)"] += dt
            elif n == 1:
                excl[allk[next(iter(live))][2]] += dt
            else:
                # PDL: a successor is resident (waiting at griddepcontrol.wait) while its predecessor runs: charge the
                # kernel that started FIRST (the one doing the work)
                first = min(live, key=lambda j: allk[j][0])
                excl[allk[first][2]] += dt
        last = t_
        if kind == 1:
            live.add(i)
        else:
            live.discard(i)
    tot = sum(excl.values())
    print(f"  in flight: 0 kernels {hist[0] / 1e6:.3f} ms, 1 kernel {hist[1] / 1e6:.3f} ms, 2 {hist[2] / 1e6:.3f} ms, >= 3 {hist[3] / 1e6:.3f} ms")
    cnt = defaultdict(int)
    for _, _, name in allk:
        cnt[name] += 1
    for name, ns in sorted(excl.items(), key=lambda kv: -kv[1])[:34]:
        print(f"    {ns / 1e6:7.3f} ms {100 * ns / tot:5.1f}%  n={cnt.get(name, 0):4d}  {name[:110]}")


run(lambda: model.apply_model(x, tt, cond), f"UNet+control step, batch {batch} (graph replay)")
run(lambda: model.decode_first_stage_u8(z), f"VAE decode to uint8, batch {batch}")
