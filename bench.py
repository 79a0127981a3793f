#!/usr/bin/env python
"""bench.py — decoded images/s of the RDEIC relay decode (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A bench "step" is one pass of the hot path over one batch: BASELINE config[1] — 512x512, batch 8
per GPU, 5 relay steps (q_sample -> 5 x UNet+control step + posterior update -> VAE decode ->
uint8), synthetic inputs, seeded random-init weights of the SD-2.1 + adapter + VAE architecture.
Prints ONE JSON line (rank 0).  `value` is timed with the inputs resident in HBM; `e2e` is the
same decode through the public API with pinned HOST buffers (H2D of the conditioning and noise,
D2H of the uint8 images inside the timed region).  `--impl reference` times the CPU oracle port of
the reference path (oracle/) on the box's host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "decoded images/s @512^2 (5 relay steps)"
UNIT = "images/s"
WORKLOAD = "512x512 batch 8 per GPU, 5 relay steps (SpacedSampler), UNet+control+VAE decode to uint8 (arithmetic type: see dtype)"
H = W = 512
BATCH = 8
RELAY_STEPS = 5
HINT_C, CTX_DIM = 256, 1024
WEIGHT_SEED = 231


def make_inputs(batch: int, h: int, w: int, seed_off: int = 0):
    g = lambda s: torch.Generator().manual_seed(s + seed_off)
    c_latent = torch.randn(batch, 4, h, w, generator=g(7))
    hint = torch.randn(batch, HINT_C, h, w, generator=g(8))
    ctx = torch.randn(batch, 77, CTX_DIM, generator=g(9))
    gn = g(231)
    noises = [torch.randn(batch, 4, h, w, generator=gn) for _ in range(RELAY_STEPS + 1)]
    return c_latent, hint, ctx, noises


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
def bench_config(world: int) -> dict:
    """The `config` object of the JSON line: identical for both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": f"dp{world}",
            "l2": "no flush needed: per-step working set (1.9 GB bf16 weights + >1 GB activations) exceeds the 126 MB L2"}


def cpu_decode_one_image(sd_cpu, c_latent, hint, ctx, noises):
    """One image through the oracle (plain PyTorch fp32 restatement of the reference modules): q_sample at t = 299,
    RELAY_STEPS SpacedSampler steps over UNet + control adapter, VAE decode, uint8 -- the whole path of one unit."""
    from oracle import nn as onn
    from oracle import sampler as osamp

    kw = dict(model_channels=320, base_d_head=64, ctrl_d_head=16)
    x_T = osamp.q_sample(c_latent, 299, noises[0])
    apply_model = lambda x, t: onn.noise_estimator_forward(sd_cpu, x, hint, t, ctx, **kw)
    z = osamp.spaced_sample(apply_model, x_T, RELAY_STEPS, noises[1:])
    return onn.to_uint8(onn.vae_decode(sd_cpu, z))


def cpu_reference_sample(sd_cpu, steps_to_time: int, warmup: int):
    """Time the oracle on a BOUNDED sample of the workload: each step decodes ONE 512x512 image of the batch of 8,
    completely (5 relay steps + VAE decode), on all host cores.  Nothing is extrapolated: `ms_per_step` is the measured
    mean of the timed steps and `value` = 1 image / that."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c_latent, hint, ctx, noises = make_inputs(1, H // 8, W // 8)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps_to_time):
            t0 = time.perf_counter()
            cpu_decode_one_image(sd_cpu, c_latent, hint, ctx, noises)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    t_img = float(np.mean(times))
    return {"value": 1.0 / t_img, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"each step = 1 of the batch's 8 images, 512x512, decoded completely (q_sample, {RELAY_STEPS} relay steps "
                      f"over UNet+control, VAE decode, uint8) by the fp32 oracle port on {cores} host threads; "
                      f"{steps_to_time} step(s) timed after {warmup} warm-up, mean {t_img:.2f} s per image",
            "s_per_image": t_img}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rdeic_b200 import configs, synthetic

    sd = synthetic.make_state_dict(configs.default_params(), seed=WEIGHT_SEED, device="cpu")
    t0 = time.perf_counter()
    res = cpu_reference_sample(sd, max(1, args.steps), max(0, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["s_per_image"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(max(1, args.gpus)),
            "reference_note": "CPU oracle port of the reference modules (the reference is pure Python with five absent "
                              "dependencies and cannot travel to the GPU box); a step is a bounded sample of the workload: "
                              "see cpu_baseline.sample",
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    real_stdout = None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL prints its version banner from C on fd 1, so fd 1 is
        # pointed at stderr for the whole run and the line is written to the saved descriptor at the end
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    from rdeic_b200 import RDEIC, build, configs, ops, parallel, synthetic
    from rdeic_b200.pipeline import relay_decode

    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    params = configs.default_params()
    spec = [(k, s) for k, s, _ in synthetic.state_dict_spec(params)]
    sd = synthetic.make_state_dict(params, seed=WEIGHT_SEED, device=dev) if rank == 0 else None
    sd = parallel.broadcast_state_dict(sd, spec, dev, src=0)          # one-time NCCL weight broadcast
    model = RDEIC.from_config({"params": params}, device=dev)
    model.load_state_dict(sd)
    sd_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd_cpu = {k: v.cpu() for k, v in sd.items()}
    del sd
    torch.cuda.empty_cache()

    h, w = H // 8, W // 8
    c_latent, hint, ctx, noises = make_inputs(BATCH, h, w, seed_off=1000 * rank)
    pin = lambda t: t.contiguous().pin_memory()
    host = {"c_latent": pin(c_latent), "hint": pin(hint), "ctx": pin(ctx), "noises": [pin(n) for n in noises]}
    d = {"c_latent": c_latent.to(dev), "hint": hint.to(dev), "ctx": ctx.to(dev), "noises": [n.to(dev) for n in noises]}
    # a second HBM-resident batch: the device-timed steps alternate between the two, so every step decodes a batch
    # whose conditioning differs from the previous step's and the step-invariant work done once per batch (cross-
    # attention K/V of the text context, NHWC bf16 hint: NoiseEstimatorEngine.prepare_cond) is inside every step
    c2, h2, x2, n2 = make_inputs(BATCH, h, w, seed_off=1000 * rank + 500)
    d_alt = {"c_latent": c2.to(dev), "hint": h2.to(dev), "ctx": x2.to(dev), "noises": [n.to(dev) for n in n2]}
    turn = [0]
    h2d = sum(t.numel() * t.element_size() for t in [host["c_latent"], host["hint"], host["ctx"], *host["noises"]])
    d2h = BATCH * H * W * 3                                   # uint8 images read back per step

    def decode_device():
        b = (d, d_alt)[turn[0] & 1]
        turn[0] += 1
        eng = model.control_model          # a new batch: nothing of the previous one's conditioning is reused
        eng._ctx_cache.clear()
        eng._hint_cache.clear()
        cond = {"c_latent": [b["c_latent"]], "c_crossattn": [b["ctx"]], "guide_hint": b["hint"]}
        return relay_decode(model, cond, RELAY_STEPS, sampler="ddpm", start_noise=b["noises"][0],
                            step_noises=b["noises"][1:])

    # e2e: the public serving call for host-resident conditioning (pipeline.decode_host_batches): every step's
    # inputs are copied host -> device from pinned memory and its uint8 images device -> host inside the timed
    # region; the copies of step i+1 / i-1 are double buffered against the decode of step i
    from rdeic_b200.pipeline import decode_host_batches

    host_batch = {"c_latent": host["c_latent"], "guide_hint": host["hint"], "c_crossattn": host["ctx"],
                  "start_noise": host["noises"][0], "step_noises": host["noises"][1:]}
    e2e_last = {}

    def decode_e2e(k):
        for out in decode_host_batches(model, (host_batch for _ in range(k)), RELAY_STEPS, sampler="ddpm"):
            e2e_last["img"] = out               # pinned host uint8 [B,H,W,3], complete when yielded

    def timed(fn, k, batched=False):
        """K steps between barrier+sync on both sides, CUDA events on the launching stream, MAX over ranks."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.LAUNCHES
        e0.record()
        if batched:
            fn(k)                       # one call that runs all k steps (and ends with the last result on the host)
        else:
            for _ in range(k):
                fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ops.LAUNCHES - n0

    for _ in range(max(args.warmup, 3)):
        decode_device()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    if args.profile_range:       # `ncu --profile-from-start off`: the launch list of exactly this region
        torch.cuda.cudart().cudaProfilerStart()
    ms_dev, launches = timed(decode_device, args.steps)
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    clk = clocks.stop() if rank == 0 else None
    decode_e2e(2)
    ms_e2e, _ = timed(decode_e2e, args.steps, batched=True)
    assert tuple(e2e_last["img"].shape) == (BATCH, H, W, 3) and d2h == e2e_last["img"].numel()

    # UNet-step ms (second half of the BASELINE metric): graph replay of one UNet+control step, B=8
    cond = {"c_latent": [d["c_latent"]], "c_crossattn": [d["ctx"]], "guide_hint": d["hint"]}
    tt = torch.full((BATCH,), 224, dtype=torch.long, device=dev)
    ms_unet, _ = timed(lambda: model.apply_model(d["noises"][0], tt, cond), 10)
    ms_vae, _ = timed(lambda: model.decode_first_stage_u8(d["c_latent"]), 3)

    # roofline of the dominant kernel (tcgen05 implicit-GEMM conv/linear): per-launch CUDA events
    # around every launch of one eager decode, on the launching stream
    model.use_cuda_graph = False
    decode_device()
    torch.cuda.synchronize()
    # Park the GPU behind a ~40 ms spin every 150 profiled launches so the host stays ahead and every
    # (event, kernel, event) triple is already queued when the GPU reaches it: otherwise each bracket
    # also times the host's ~15 us Python enqueue gap, which for 1 577 short launches is a sizeable
    # over-count.  (One long spin at the start is not enough: the launch queue holds ~1 000 entries,
    # so the host can only get that far ahead.)  The spins are outside every bracket.
    class _Parked(list):
        def append(self, item):
            if len(self) % 150 == 0:
                torch.cuda._sleep(int(0.04 * 1.9e9))
            super().append(item)

    # one stream for this pass: a bracket around a kernel that shares the GPU with the control
    # adapter's stream would time the sharing, not the kernel
    overlap, model.control_model.overlap_control = model.control_model.overlap_control, False
    ops.GEMM_PROFILE = _Parked()
    decode_device()
    torch.cuda.synchronize()
    prof, ops.GEMM_PROFILE = list(ops.GEMM_PROFILE), None
    # the same serialised (one stream, no graph) decode without brackets: the denominator of the kernel's
    # share, like for like with the serialised ncu launch list under profiles/
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    decode_device()
    s1.record()
    torch.cuda.synchronize()
    ms_serial = s0.elapsed_time(s1)
    model.control_model.overlap_control = overlap
    model.use_cuda_graph = True
    g_ms = sum(a.elapsed_time(b) for a, b, _ in prof)
    g_fl = sum(f for _, _, f in prof)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved_tf = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    # DRAM bytes per launch of the same kernel: only from an ncu capture of THIS command (`--traffic-from file`, the
    # kernel summary scripts/kernel_summary.py writes from `ncu ... python bench.py --profile-range`); null otherwise
    traffic = None
    if args.traffic_from and Path(args.traffic_from).exists():
        traffic = json.loads(Path(args.traffic_from).read_text()).get("conv_gemm_dram_bytes_per_launch")
    roofline = {"kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv3x3/conv1x1/linear)", "bound": "tensor",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": traffic, "launches_per_step": len(prof),
                "flops_per_launch": g_fl / max(len(prof), 1), "ms_per_launch": g_ms / max(len(prof), 1),
                "share_of_step": g_ms / ms_serial, "share_basis": "one-stream eager decode of the same batch "
                f"({ms_serial:.1f} ms; the graphed, two-stream step is ms_per_step)",
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"}

    if rank == 0:
        n_img = BATCH * world * args.steps
        line = {"metric": METRIC, "value": n_img / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": bench_config(world),
                "arm_notes": {"sampler": "SpacedSampler fixed_small, CUDA-graphed UNet step",
                              "inputs": "two HBM-resident batches alternate step to step: the once-per-batch conditioning work "
                                        "(text K/V projections, NHWC bf16 hint) is inside every timed step; e2e: "
                                        "pipeline.decode_host_batches, H2D/D2H double buffered against the decode"},
                "unet_step_ms": ms_unet / 10, "vae_decode_ms": ms_vae / 3,
                "e2e": {"value": n_img / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline}
        if sd_cpu is not None:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_sample(sd_cpu, 2, 0).items() if k != "s_per_image"}
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# the other BASELINE.json configs (strong scaling; not the headline line the driver records)
# ---------------------------------------------------------------------------------------------
def _entropy_frontend_bitexact(dev) -> bool:
    """BASELINE config 5's check: checkerboard split / merge / squeeze and the VQ look-up on the GPU against the
    outputs of the reference's own utils/ckbd.py and VectorQuantiser (tests/golden/entropy_ref.npz), bit for bit."""
    from rdeic_b200 import ckbd, ops

    g = np.load(ROOT / "tests" / "golden" / "entropy_ref.npz")
    # the seeded inputs the goldens were made from (tests/golden/make_golden.py::entropy_inputs)
    gen = torch.Generator().manual_seed(41)
    y = torch.randn(2, 8, 6, 10, generator=gen) * 6
    y.view(-1)[3] = -0.0
    K, D = 512, 256
    cb = (torch.rand(K, D, generator=gen) * 2 - 1) / K
    cb[7] = cb[300]
    pick = torch.randint(0, K, (2 * 3 * 5,), generator=gen)
    pick[0] = 300
    z = cb[pick] + 1e-5 * torch.randn(pick.numel(), D, generator=gen)
    z[0] = cb[300]
    z = z.reshape(2, 3, 5, D).permute(0, 3, 1, 2).contiguous()
    bits = lambda t: t.cpu().numpy().view(np.uint32)
    yc = y.to(dev)
    a, n = ckbd.ckbd_split(yc)
    ok = np.array_equal(bits(a), g["anchor"].view(np.uint32)) and np.array_equal(bits(n), g["nonanchor"].view(np.uint32))
    ok &= np.array_equal(bits(ckbd.ckbd_merge(a, n)), g["merge"].view(np.uint32))
    sa, sn = ckbd.ckbd_anchor_sequeeze(yc), ckbd.ckbd_nonanchor_sequeeze(yc)
    ok &= np.array_equal(bits(sa), g["anchor_sq"].view(np.uint32)) and np.array_equal(bits(sn), g["nonanchor_sq"].view(np.uint32))
    ok &= np.array_equal(bits(ckbd.ckbd_anchor_unsequeeze(sa)), g["anchor_unsq"].view(np.uint32))
    zq, idx = ops.vq_quant(z.to(dev), cb.to(dev))
    ok &= np.array_equal(idx.cpu().numpy(), g["vq_idx"]) and np.array_equal(bits(zq), g["vq_zq"].view(np.uint32))
    # quantise / dequantise round trip on a large tensor: symbols are integers, x_hat - means is integral
    x = torch.randn(1 << 20, generator=gen).to(dev) * 6
    mu = torch.randn(1 << 20, generator=gen).to(dev) * 2
    sym = ops.quantize_symbols(x, mu)
    ok &= bool(torch.equal(sym, torch.round(x - mu).to(torch.int32)))
    ok &= bool(torch.equal(ops.dequantize(sym, mu), sym.float() + mu))
    return bool(ok)


def run_other_config(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    real_stdout = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    from rdeic_b200 import RDEIC, build, configs, ops, parallel, synthetic
    from rdeic_b200.pipeline import relay_decode

    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    params = configs.default_params()
    spec = [(k, s) for k, s, _ in synthetic.state_dict_spec(params)]
    sd = synthetic.make_state_dict(params, seed=WEIGHT_SEED, device=dev) if rank == 0 else None
    sd = parallel.broadcast_state_dict(sd, spec, dev, src=0)
    model = RDEIC.from_config({"params": params}, device=dev).load_state_dict(sd)
    del sd
    torch.cuda.empty_cache()

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.LAUNCHES
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ops.LAUNCHES - n0

    cfg = args.config
    extra = {}
    if cfg in ("c3", "c5"):
        Hh, Ww, GB = (512, 768, 64) if cfg == "c3" else (512, 512, 32)
        lo, hi = parallel.shard_range(GB, rank, world)
        counts = [parallel.shard_range(GB, r, world)[1] - parallel.shard_range(GB, r, world)[0] for r in range(world)]
        c_latent, hint, ctx, noises = make_inputs(GB, Hh // 8, Ww // 8)
        while len(noises) < 11:                                   # the 10-step sweep of c5
            noises = noises + noises[1:]
        sl = lambda t: t[lo:hi].contiguous()
        host = {"c_latent": sl(c_latent).pin_memory(), "hint": sl(hint).pin_memory(), "ctx": sl(ctx).pin_memory(),
                "noises": [sl(n).pin_memory() for n in noises[:11]]}
        d = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in host.items()}
        out_host = torch.empty((GB, Hh, Ww, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None

        def decode(src, steps, to_host=False):
            if src is host:                                       # e2e: this step's inputs cross PCIe first
                b = {k: ([t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v.to(dev, non_blocking=True))
                     for k, v in host.items() if k != "noises"}
                b["noises"] = [t.to(dev, non_blocking=True) for t in host["noises"][:steps + 1]]
            else:
                b = src
            cond = {"c_latent": [b["c_latent"]], "c_crossattn": [b["ctx"]], "guide_hint": b["hint"]}
            imgs = relay_decode(model, cond, steps, sampler="ddpm", start_noise=b["noises"][0], step_noises=b["noises"][1:steps + 1])
            allimg = parallel.gather_images(imgs, counts, dst=0)          # NCCL gather of the uint8 images, inside the timed region
            if to_host and rank == 0:
                out_host.copy_(allimg, non_blocking=True)
            return allimg

        for _ in range(max(args.warmup, 3)):
            decode(d, RELAY_STEPS)
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        ms_dev, launches = timed(lambda: decode(d, RELAY_STEPS), args.steps)
        clk = clocks.stop() if rank == 0 else None
        ms_e2e, _ = timed(lambda: decode(host, RELAY_STEPS, to_host=True), args.steps)
        h2d = sum(t.numel() * 4 for t in [host["c_latent"], host["hint"], host["ctx"], *host["noises"][:RELAY_STEPS + 1]])
        if cfg == "c5":
            sweep = {}
            for st in (2, 5, 10):
                decode(d, st)
                ms, _ = timed(lambda: decode(d, st), max(2, args.steps // 2))
                sweep[str(st)] = GB * max(2, args.steps // 2) / (ms * 1e-3)
            extra = {"images_per_s_by_relay_steps": sweep, "entropy_frontend_bitexact": _entropy_frontend_bitexact(dev)}
        value, e2e_v, unit_imgs = GB * args.steps / (ms_dev * 1e-3), GB * args.steps / (ms_e2e * 1e-3), GB
        workload = (f"{Ww}x{Hh} global batch {GB} split over {world} GPU(s) ({hi - lo} per GPU), 5 relay steps (SpacedSampler), bf16, "
                    "UNet+control+VAE decode to uint8, NCCL gather of the uint8 images to rank 0 inside the timed region")
        metric = f"decoded images/s @{Ww}x{Hh} (5 relay steps)"
        d2h = GB * Hh * Ww * 3
    else:   # c4: one 2048x1365 image (padded to 2048x1408 -> latent 176x256) as overlapping latent tiles
        Hl, Wl, OV = 1408 // 8, 2048 // 8, 16
        g = torch.Generator().manual_seed(4)
        host = {"c_latent": torch.randn(1, 4, Hl, Wl, generator=g).pin_memory(), "ctx": torch.randn(1, 77, CTX_DIM, generator=g).pin_memory(),
                "hint": torch.randn(1, HINT_C, Hl, Wl, generator=g).pin_memory()}
        d = {k: v.to(dev) for k, v in host.items()}
        plan = parallel.plan_tiles_balanced(Hl, Wl, world, overlap=OV, max_tile_area=args.c4_max_tile_area)
        out_host = torch.empty((Hl * 8, Wl * 8, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None

        def decode(src, to_host=False):
            b = {k: v.to(dev, non_blocking=True) for k, v in host.items()} if src is host else src
            cond = {"c_latent": [b["c_latent"]], "c_crossattn": [b["ctx"]], "guide_hint": b["hint"]}
            img = parallel.decode_tiled_u8(lambda c, idx: relay_decode(model, c, RELAY_STEPS), cond, plan, overlap=OV)
            if to_host and rank == 0:
                out_host.copy_(img, non_blocking=True)
            return img

        for _ in range(max(args.warmup, 3)):
            img = decode(d)
        if rank == 0:
            assert tuple(img.shape) == (Hl * 8, Wl * 8, 3)
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        ms_dev, launches = timed(lambda: decode(d), args.steps)
        clk = clocks.stop() if rank == 0 else None
        ms_e2e, _ = timed(lambda: decode(host, to_host=True), args.steps)
        h2d = sum(t.numel() * 4 for t in host.values())
        value, e2e_v, unit_imgs = args.steps / (ms_dev * 1e-3), args.steps / (ms_e2e * 1e-3), 1
        th, tw = plan[0][2], plan[0][3]
        extra = {"tiles": len(plan), "tile_latent": [th, tw], "overlap_latent": OV, "tiles_per_gpu": len(plan) // world,
                 "mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30}
        workload = (f"one 2048x1365 image (padded 2048x1408, latent {Hl}x{Wl}) as {len(plan)} overlapping {th}x{tw} latent tiles dealt "
                    f"over {world} GPU(s), 5 relay steps, uint8 tiles gathered to rank 0 and blended by one kernel")
        metric = "decoded images/s @2048x1365 tiled (5 relay steps)"
        d2h = Hl * 8 * Wl * 8 * 3
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "baseline_config": cfg, "global_batch": unit_imgs, "parallelism": f"dp{world}",
                           "l2": "no flush needed: per-step working set exceeds the 126 MB L2"},
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clk, **extra}
        if real_stdout is not None:
            sys.stdout.flush()
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# the learned compressor either side of the decode (not the headline line)
# ---------------------------------------------------------------------------------------------
def run_compressor(args):
    """`--compressor B`: compress / decompress of B 512x512 images through `rdeic_b200.compression.
    Compression` (feature map [B,512,64,64] -> y [B,256,32,32], z [B,256,8,8]); byte coders replaced by
    an in-memory replay (rANS / torchac are host libraries outside this path); at B = 1 the CPU oracle
    (reference algorithm, torch fp32, all host cores) is timed beside it on the same image."""
    from rdeic_b200 import build, configs, ops, synthetic
    from rdeic_b200.compression import Compression

    build.build()
    B = args.compressor
    dev = torch.device("cuda:0")
    pp = configs.default_params()["preprocess_config"]["params"]
    sd = synthetic.make_compression_state_dict(pp, seed=232)

    class Loop:
        accepts_arrays = True       # int32 numpy views of the pinned hand-off buffers, no Python lists

        def __init__(self):
            self.symbols, self.pos = [], 0

        def encode_with_indexes(self, symbols, indexes, *a):
            self.symbols = symbols.copy()

        def flush(self):
            return b""

        def set_stream(self, s):
            self.pos = 0

        def decode_stream(self, indexes, *a):
            n = len(indexes)
            out = self.symbols[self.pos:self.pos + n]
            self.pos += n
            return out

    class Hyp:
        def compress(self, idx):
            return idx

        def decompress(self, s, shape):
            return s

    loop = Loop()
    m = Compression(device=dev, rans_encoder=lambda: loop, rans_decoder=lambda: loop, hyper_latent_coder=Hyp(),
                    precision=args.compressor_precision, **pp)
    m.load_state_dict(sd)
    x = torch.randn(B, pp["in_nc"], 64, 64, generator=torch.Generator().manual_seed(3))

    def timed(fn, n):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3, r

    m.use_cuda_graph = False
    ops.LAUNCHES = 0
    out = m.compress(x)
    torch.cuda.synchronize()
    launches_c = ops.LAUNCHES
    ops.LAUNCHES = 0
    m.decompress(out["strings"], out["shape"])
    torch.cuda.synchronize()
    launches_d = ops.LAUNCHES
    ms_c_eager, _ = timed(lambda: m.compress(x), 5)
    ms_d_eager, _ = timed(lambda: m.decompress(out["strings"], out["shape"]), 5)
    m.use_cuda_graph = True
    for _ in range(2):                     # a shape gets its CUDA-graph plan the second time it is seen
        out = m.compress(x)
        m.decompress(out["strings"], out["shape"])
    ms_c, out = timed(lambda: m.compress(x), 10)
    ms_d, _ = timed(lambda: m.decompress(out["strings"], out["shape"]), 10)
    res = {"workload": f"512x512 image, batch {B}: feature map [B,512,64,64] -> y [B,256,32,32], z [B,256,8,8]",
           "precision": args.compressor_precision,
           "compress_ms": ms_c, "decompress_ms": ms_d, "compress_ms_no_graph": ms_c_eager,
           "decompress_ms_no_graph": ms_d_eager, "kernel_launches": {"compress": launches_c, "decompress": launches_d},
           "symbols_per_image": len(loop.symbols) // B}
    if B == 1:
        from oracle import compression_nets as ocn

        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        r = ocn.compress(sd, x, pp["slice_ch"])
        t1 = time.perf_counter()
        ocn.decompress(sd, r["z_idx"].numpy(), r["symbols"], r["indexes"], pp["slice_ch"])
        t2 = time.perf_counter()
        res["cpu_baseline"] = {"compress_ms": (t1 - t0) * 1e3, "decompress_ms": (t2 - t1) * 1e3, "cores": os.cpu_count(),
                               "kind": "port", "sample": "one 512x512 image"}
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed device region (for ncu --profile-from-start off)")
    ap.add_argument("--traffic-from", default=None, metavar="JSON",
                    help="kernel summary of an ncu capture of this same command: fills roofline.traffic (else null)")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json config: c2 (default, the headline: 512^2 batch 8 per GPU, weak scaling), c3 (768x512 "
                         "global batch 64 split over the GPUs, strong scaling, uint8 gather timed), c4 (one 2048x1365 image as "
                         "latent tiles dealt over the GPUs), c5 (512^2 global batch 32, relay steps 2/5/10 + entropy front-end "
                         "bit-exactness against the reference goldens)")
    ap.add_argument("--compressor", type=int, default=0, metavar="B",
                    help="measure the learned compressor (SURVEY 8f rows) on B 512x512 images instead of the decode")
    ap.add_argument("--c4-max-tile-area", type=int, default=96 * 96, help="config c4: largest latent tile area the planner may pick")
    ap.add_argument("--compressor-precision", default="mixed", choices=["mixed", "bf16", "fp32"],
                    help="mixed (default): entropy-parameter nets on the fp32 kernels (streams exchangeable with the reference)")
    args = ap.parse_args()
    if args.compressor:
        run_compressor(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.config != "c2":
        run_other_config(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
