"""BASELINE.json config 4: ONE 2048x1365 image (padded to 2048x1408 -> latent 176x256), 5 relay steps,
decoded as overlapping 64x64 latent tiles dealt round-robin across the ranks (each rank decodes its
tiles as one batch), blended on rank 0.  Tiling is new behaviour (the reference never tiles); parity
of the tiled path is tested per tile against the oracle (tests/test_gpu_engine.py).
    python scripts/run_c4_tiled.py            # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/run_c4_tiled.py"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rdeic_b200 import RDEIC, configs, parallel, synthetic  # noqa: E402
from rdeic_b200.pipeline import relay_decode  # noqa: E402

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
params = configs.default_params()
spec = [(k, s) for k, s, _ in synthetic.state_dict_spec(params)]
sd = synthetic.make_state_dict(params, seed=231, device=dev) if rank == 0 else None
model = RDEIC.from_config({"params": params}, device=dev).load_state_dict(parallel.broadcast_state_dict(sd, spec, dev, src=0))
del sd
H, W, TILE, OV, STEPS = 1408 // 8, 2048 // 8, 64, 16, 5
g = torch.Generator().manual_seed(4)
cond = {"c_latent": [torch.randn(1, 4, H, W, generator=g).to(dev)], "c_crossattn": [torch.randn(1, 77, 1024, generator=g).to(dev)],
        "guide_hint": torch.randn(1, 256, H, W, generator=g).to(dev)}
plan = parallel.plan_tiles(H, W, TILE, OV)


def decode(c, idx):
    return relay_decode(model, c, STEPS, as_uint8=False)


def once():
    return parallel.decode_tiled(decode, cond, tile=TILE, overlap=OV, batched=True)


for _ in range(2):
    img = once()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
e0.record()
for _ in range(K):
    img = once()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    assert tuple(img.shape) == (3, 1408, 2048) and torch.isfinite(img).all()
    print(json.dumps({"config": "C4 2048x1365 (padded 2048x1408), 5 relay steps, tiled", "n_gpus": world, "tiles": len(plan),
                      "tile_latent": TILE, "overlap_latent": OV, "tiles_per_rank_max": -(-len(plan) // world),
                      "ms_per_image": float(ms.item()), "images_per_s": 1e3 / float(ms.item()),
                      "mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30}))
if world > 1:
    dist.destroy_process_group()
