"""Roofline of the bandwidth-bound kernels on tensors far larger than L2 (>= 2^26 elements), plus
their in-situ latency on decode-sized tensors.  Prints one JSON object.
achieved GB/s = algorithmic bytes (each operand read once, each result written once) / CUDA-event time."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import ops  # noqa: E402
from rdeic_b200.ckbd import get_scale_table  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text()) \
    if (Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").exists() else {}
PEAK = float(peaks.get("hbm_gbs", 6650.0))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


res = {}


def rec(name, nbytes, fn, small_fn=None):
    t = timeit(fn)
    d = {"GBps": nbytes / t / 1e9, "frac_of_measured_hbm": nbytes / t / 1e9 / PEAK, "ms": t * 1e3, "bytes": nbytes}
    if small_fn is not None:
        d["in_situ_us"] = timeit(small_fn, 30) * 1e6
    res[name] = d


g = torch.Generator(device=dev).manual_seed(0)
# --- entropy front end: [B,C,H,W] fp32, 2^26 elements ---------------------------------------
B, C, Hh, Ww = 16, 64, 256, 256
n = B * C * Hh * Ww
y = torch.randn(B, C, Hh, Ww, generator=g, device=dev) * 6
mu = torch.randn(B, C, Hh, Ww, generator=g, device=dev) * 2
sc = torch.exp(torch.rand(B, C, Hh, Ww, generator=g, device=dev) * 8 - 3)
table = get_scale_table().to(dev)
ys = torch.randn(1, 64, 32, 32, generator=g, device=dev)          # decode-sized slice @512^2
rec("ckbd_split", 12 * n, lambda: ops.ckbd_split(y), lambda: ops.ckbd_split(ys))
rec("ckbd_merge", 12 * n, lambda: ops.ckbd_merge(y, mu), lambda: ops.ckbd_merge(ys, ys))
rec("ckbd_squeeze", 6 * n, lambda: ops.ckbd_squeeze(y, 0), lambda: ops.ckbd_squeeze(ys, 0))
sq = ops.ckbd_squeeze(y, 0)
rec("ckbd_unsqueeze", 6 * n, lambda: ops.ckbd_unsqueeze(sq, 0))
rec("quantize_symbols", 12 * n, lambda: ops.quantize_symbols(y, mu), lambda: ops.quantize_symbols(ys, ys))
sym = ops.quantize_symbols(y, mu)
rec("dequantize", 12 * n, lambda: ops.dequantize(sym, mu))
rec("build_indexes", 8 * n, lambda: ops.build_indexes(sc, table, 0.11), lambda: ops.build_indexes(ys.abs(), table, 0.11))
# fused phases touch only half the positions of each input: y,scales,means read (3 * n/2 * 4) ...
rec("ckbd_encode_phase(fused)", 4 * (3 * n // 2 + 2 * (n // 2) + n), lambda: ops.ckbd_encode_phase(y, sc, mu, table, 0.11, 0),
    lambda: ops.ckbd_encode_phase(ys, ys.abs(), ys, table, 0.11, 0))
rec("ckbd_squeeze_indexes(fused)", 4 * (2 * n // 2 + 2 * (n // 2)), lambda: ops.ckbd_squeeze_indexes(sc, mu, table, 0.11, 0))
del sq, sym
# --- sampler update: 4 fp32 tensors ------------------------------------------------------------
m = 1 << 26
x, e, nz = (torch.randn(m, generator=g, device=dev) for _ in range(3))
xs = torch.randn(8, 4, 64, 64, generator=g, device=dev)
rec("relay_update", 16 * m, lambda: ops.relay_update(x, e, nz, 1.2, 0.8, 0.4, 0.6, 0.3),
    lambda: ops.relay_update(xs, xs, xs, 1.2, 0.8, 0.4, 0.6, 0.3))
rec("q_sample", 12 * m, lambda: ops.q_sample(x, nz, 0.77, 0.64))
del x, e, nz
# --- normalisation: NHWC bf16, read twice? no: algorithmic = read once + write once ------------
gn_in = torch.randn(8, 512, 512, 128, generator=g, device=dev).bfloat16()     # VAE top level, 2^28 elements
gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
small = torch.randn(8, 64, 64, 320, generator=g, device=dev).bfloat16()
g320, b320 = torch.ones(320, device=dev), torch.zeros(320, device=dev)
rec("groupnorm_silu[8,512,512,128]bf16", 4 * gn_in.numel(), lambda: ops.groupnorm(gn_in, gam, bet, 32, 1e-6, True),
    lambda: ops.groupnorm(small, g320, b320, 32, 1e-5, True))
# the decode path: statistics come from the producing GEMM's epilogue (slab table), GroupNorm = fold + apply
w128 = ops.pack_conv_weight(torch.randn(128, 128, 3, 3, generator=g, device=dev) / 34)
_, st128 = ops.conv_gemm(gn_in, w128, 128, 9, stats=True)
rec("groupnorm_silu[8,512,512,128]bf16, statistics from the producer's epilogue", 4 * gn_in.numel(),
    lambda: ops.groupnorm(gn_in, gam, bet, 32, 1e-6, True, stats1=st128))
del st128
f32s = small.float()
rec("groupnorm_silu[8,64,64,320]f32->bf16(in-situ size)", 6 * small.numel(), lambda: ops.groupnorm(f32s, g320, b320, 32, 1e-5, True))
ln_in = torch.randn(1 << 18, 320, generator=g, device=dev).bfloat16()
rec("layernorm[2^18,320]bf16", 4 * ln_in.numel(), lambda: ops.layernorm(ln_in, g320, b320),
    lambda: ops.layernorm(small.view(-1, 320), g320, b320))
ln32 = ln_in.float()
rec("layernorm[2^18,320]f32->bf16 (the UNet token stream)", 6 * ln_in.numel(), lambda: ops.layernorm(ln32, g320, b320),
    lambda: ops.layernorm(f32s.view(-1, 320), g320, b320))
del ln32
up_in = torch.randn(8, 256, 256, 256, generator=g, device=dev).bfloat16()
rec("upsample2x[8,256,256,256]", 2 * up_in.numel() * 5, lambda: ops.upsample2x(up_in))
print(json.dumps({"peak_gbs_measured": PEAK, "kernels": res}))
