"""Print the rel-L2 of every compressor conv stack against the reference goldens (diagnostic)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import rel_l2  # noqa: E402
from oracle import compression_nets as ocn  # noqa: E402
from rdeic_b200 import configs, synthetic  # noqa: E402
from rdeic_b200.compression import Compression  # noqa: E402

cuda = torch.device("cuda:0")
for tag in ("small", "full"):
    params = configs.small_params() if tag == "small" else configs.default_params()
    pp = params["preprocess_config"]["params"]
    sd = synthetic.make_compression_state_dict(pp, seed=232)
    g = np.load(ROOT / "tests" / "golden" / f"{tag}_compression.npz")
    m = Compression(device=cuda, **pp).load_state_dict(sd)
    y, z = m.analysis(torch.from_numpy(g["x"]))
    print(tag, "y", rel_l2(y.cpu(), g["y"]), "z(own y)", rel_l2(z.cpu(), g["z"]))
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(cuda).bfloat16()
    zz = m.hyper_enc(nhwc(torch.from_numpy(g["y"])), out_f32=True).permute(0, 3, 1, 2)
    print(tag, "z(ref y)", rel_l2(zz.cpu(), g["z"]))
    z_q = m.quantize.get_codebook_entry(torch.from_numpy(g["z_idx"]))
    hyper = m._hyper_params(z_q).float().permute(0, 3, 1, 2)
    print(tag, "hyper", rel_l2(hyper.cpu(), g["hyper_params"]))
    c_ref, gh_ref, y_hat, _ = ocn.decompress(sd, g["z_idx"], g["symbols"].tolist(), g["indexes"].tolist(), pp["slice_ch"])
    c, gh = m._synthesis(y_hat.to(cuda))
    print(tag, "guide_hint", rel_l2(gh.cpu(), g["guide_hint"]), "c_latent", rel_l2(c.cpu(), g["c_latent"]))
    # per-block drift through g_s
    x = nhwc(y_hat)
    import torch.nn.functional as F
    xr = y_hat
    P = "preprocess_model.decoder.g_s."
    for i, blk in enumerate(m.decoder.blocks):
        x = blk(x)
        if i == 0:
            xr = ocn._conv(sd, P + "0", xr)
        elif i == 4:
            xr = ocn.residual_block_upsample(sd, P + "4", xr)
        else:
            xr = ocn.residual_block(sd, P + str(i), xr)
        print("   g_s block", i, rel_l2(x.float().permute(0, 3, 1, 2).cpu(), xr))
