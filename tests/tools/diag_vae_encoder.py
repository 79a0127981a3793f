"""Per-stage rel-L2 of the VAE encoder engine against the CPU oracle (diagnostic)."""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import rel_l2  # noqa: E402
from oracle import nn as onn  # noqa: E402
from rdeic_b200 import configs, ops, synthetic  # noqa: E402
from rdeic_b200.engine import Act, VAEEncoderEngine  # noqa: E402

dev = torch.device("cuda:0")
for tag in ("small", "full"):
    params = configs.small_params() if tag == "small" else configs.default_params()
    sd = synthetic.make_state_dict(params, seed=231, encoder=True)
    g = np.load(ROOT / "tests" / "golden" / f"{tag}_vae_encode.npz")
    e = VAEEncoderEngine(sd, device=dev)
    x = torch.from_numpy(g["x"])
    E = "first_stage_model.encoder"
    hr = onn._conv(sd, E + ".conv_in", x)
    a = Act(*ops.conv_gemm(ops.nchw_to_nhwc_bf16(x.to(dev), ldc=8), e.conv_in.w, e.conv_in.n_out, 9, bias=e.conv_in.b, dual=True))
    cmp = lambda name, t, r: print(tag, name, "%.2e" % rel_l2(t.float().permute(0, 3, 1, 2).cpu(), r))
    cmp("conv_in", a.f, hr)
    for lvl, (blocks, down) in enumerate(e.levels):
        for i, b in enumerate(blocks):
            a = e._res32(b, a)
            hr = onn._vae_resnet(sd, f"{E}.down.{lvl}.block.{i}", hr)
            cmp(f"down{lvl}.block{i}", a.f, hr)
        if down is not None:
            B, H, W, C = a.h.shape
            col = ops.im2col_3x3_s2(a.h, pad_lo=0)
            of, oh = ops.linear(col, down.w, down.n_out, bias=down.b, dual=True)
            a = Act(of.view(B, H // 2, W // 2, down.n_out), oh.view(B, H // 2, W // 2, down.n_out))
            w, bb = sd[f"{E}.down.{lvl}.downsample.conv.weight"], sd[f"{E}.down.{lvl}.downsample.conv.bias"]
            hr = F.conv2d(F.pad(hr, (0, 1, 0, 1)), w, bb, stride=2)
            cmp(f"down{lvl}.downsample", a.f, hr)
    a = e._res32(e.mid1, a)
    hr = onn._vae_resnet(sd, E + ".mid.block_1", hr)
    cmp("mid1", a.f, hr)
    t = e._attn(a.h)[0]
    hr = onn._vae_attn(sd, E + ".mid.attn_1", hr)
    cmp("attn", t, hr)
    a = e._res32(e.mid2, Act(None, t))
    hr = onn._vae_resnet(sd, E + ".mid.block_2", hr)
    cmp("mid2", a.f, hr)
    c = ops.groupnorm(a.f, e.norm_out.g, e.norm_out.b, 32, 1e-6, True)
    cmp("c", c, torch.from_numpy(g["c"]))
