"""SURVEY §8(f) ranks 1-3 measurement: the learned compressor's compress / decompress at the
BASELINE config[1] image size (512x512 -> 64x64x512 feature map, y 32x32x256, z 8x8), coders replaced
by an in-memory replay (the rANS / torchac byte coders are host libraries outside this path), with
the CPU oracle (reference algorithm, torch fp32) timed beside it on a bounded sample.
Usage: python scripts/bench_compression.py [batch] [out.json]"""
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from rdeic_b200 import configs, ops, synthetic  # noqa: E402
from rdeic_b200.compression import Compression  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
out_path = sys.argv[2] if len(sys.argv) > 2 else None
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
m = Compression(device=dev, rans_encoder=lambda: loop, rans_decoder=lambda: loop, hyper_latent_coder=Hyp(), **pp)
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
m.use_cuda_graph = True
ms_c, out = timed(lambda: m.compress(x), 10)
ms_d, _ = timed(lambda: m.decompress(out["strings"], out["shape"]), 10)
m.use_cuda_graph = False
ms_c_eager, _ = timed(lambda: m.compress(x), 5)
ms_d_eager, _ = timed(lambda: m.decompress(out["strings"], out["shape"]), 5)
m.use_cuda_graph = True
# GPU-only part of decompress: hyper decoder + synthesis (no host hand-off in between)
z_q = m.quantize.get_codebook_entry(out["strings"][1][0].long())
ms_h, hyper = timed(lambda: m._hyper_params(z_q), 20)
y_hat = torch.randn(B, pp["M"], 32, 32, device=dev)
ms_s, _ = timed(lambda: m._synthesis(y_hat), 20)

res = {"workload": f"512x512 image, batch {B}: feature map [B,512,64,64] -> y [B,256,32,32], z [B,256,8,8]",
       "compress_ms": ms_c, "decompress_ms": ms_d, "compress_ms_no_graph": ms_c_eager,
       "decompress_ms_no_graph": ms_d_eager, "hyper_decoder_ms": ms_h, "synthesis_ms": ms_s,
       "kernel_launches": {"compress": launches_c, "decompress": launches_d},
       "symbols_per_image": len(loop.symbols) // B}
if B == 1:
    from oracle import compression_nets as ocn

    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    r = ocn.compress(sd, x, pp["slice_ch"])
    t1 = time.perf_counter()
    ocn.decompress(sd, r["z_idx"].numpy(), r["symbols"], r["indexes"], pp["slice_ch"])
    t2 = time.perf_counter()
    res["cpu_oracle"] = {"compress_ms": (t1 - t0) * 1e3, "decompress_ms": (t2 - t1) * 1e3, "cores": os.cpu_count(),
                         "kind": "port", "sample": "one 512x512 image"}
print(json.dumps(res))
if out_path:
    Path(out_path).write_text(json.dumps(res, indent=1))
