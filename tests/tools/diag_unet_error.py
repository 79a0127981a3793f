"""Where does the bf16 step lose accuracy?  Run the base UNet (unconditional branch) block by block in the
bf16 throughput engine and in the fp32 kernel mode and print the rel-L2 of the residual stream after
every block (full width, 256x256)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import rel_l2  # noqa: E402
from rdeic_b200 import configs, ops, synthetic  # noqa: E402
from rdeic_b200.engine import Act, NoiseEstimatorEngine, _Ctx  # noqa: E402
from rdeic_b200.engine_f32 import NoiseEstimatorF32  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 231
dev = torch.device("cuda:0")
params = configs.default_params()
up, cp = params["unet_config"]["params"], params["control_stage_config"]["params"]
sd = synthetic.make_state_dict(params, seed=seed, device=dev)
e16 = NoiseEstimatorEngine(sd, up, cp, device=dev)
e32 = NoiseEstimatorF32(sd, up, cp, device=dev)
g = torch.Generator(device=dev).manual_seed(seed + 100)
B, h = 1, (int(sys.argv[2]) if len(sys.argv) > 2 else 32)
x = torch.randn(B, 4, h, h, generator=g, device=dev)
ctx = torch.randn(B, 77, 1024, generator=g, device=dev)
t = torch.full((B,), 224, dtype=torch.long, device=dev)
with torch.no_grad():
    # bf16 engine pieces
    kvb, kvc, _ = e16.prepare_cond(ctx, None)
    hb, t_emb = e16._step_inputs(x, t)
    cb = _Ctx(e16._time_rows(e16.base, t_emb), kvb)
    # fp32 pieces
    B_ = "model.diffusion_model"
    h32 = x.permute(0, 2, 3, 1).contiguous()
    te = torch.empty((B, e32.model_channels), device=dev)
    ops.check(ops._lib.load().rdeic_timestep_embedding_f32(t.data_ptr(), te.data_ptr(), B, e32.model_channels, 10000.0,
                                                           torch.cuda.current_stream().cuda_stream), "te")
    eb = e32._emb(B_, te)
    hs16, hs32 = [], []
    cmp = lambda name, a, b: print(f"{name:22s} {rel_l2(a.f.cpu() if a.f is not None else a.h.float().cpu(), b.cpu()):.3e}")
    for i, blk in enumerate(e16.base.input_blocks):
        hb = e16._run_block(blk, hb, None, cb)
        h32 = e32.block(f"{B_}.input_blocks.{i}", h32, None, eb, ctx, e32.base_d_head, False)
        hs16.append(hb); hs32.append(h32)
        cmp(f"input_blocks.{i}", hb, h32)
    hb = e16._run_block(e16.base.middle, hb, None, cb)
    h32 = e32.block(f"{B_}.middle_block", h32, None, eb, ctx, e32.base_d_head, False)
    cmp("middle_block", hb, h32)
    for i, blk in enumerate(e16.base.output_blocks):
        hb = e16._run_block(blk, hb, hs16.pop(), cb)
        h32 = e32.block(f"{B_}.output_blocks.{i}", h32, hs32.pop(), eb, ctx, e32.base_d_head, False)
        cmp(f"output_blocks.{i}", hb, h32)
