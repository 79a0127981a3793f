"""Architecture hyper-parameters of the RDEIC decode path.

The drop-in consumes the reference's `configs/model/rdeic.yaml` unchanged
(`RDEIC.from_config(path)`).  Benchmarks and tests that run where the reference tree is absent
(the GPU box) use `default_params()`, which restates only the decode-relevant fields of that YAML
(reference configs/model/rdeic.yaml:3-15,26,33-67,69-91,102-111).
"""
from __future__ import annotations

import copy
from typing import Any, Dict


def default_params() -> Dict[str, Any]:
    unet = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
                use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, legacy=False)
    ctrl = dict(unet, hint_channels=256, num_head_channels=16, control_model_ratio=0.2, control_scale=1.0)
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    return dict(linear_start=0.00085, linear_end=0.0120, timesteps=1000, scale_factor=0.18215, used_timesteps=300,
                control_stage_config=dict(target="model.rdeic.NoiseEstimator", params=ctrl),
                unet_config=dict(target="ldm.modules.diffusionmodules.openaimodel.UNetModel", params=unet),
                first_stage_config=dict(target="ldm.models.autoencoder.AutoencoderKL",
                                        params=dict(embed_dim=4, ddconfig=dd)),
                preprocess_config=dict(target="model.compression.Compression",
                                       params=dict(in_nc=512, out_nc=4, N=256, M=256, slice_num=10,
                                                   slice_ch=[8, 8, 8, 8, 16, 16, 32, 32, 64, 64], codebook_size=16384)))


def small_params() -> Dict[str, Any]:
    """Reduced-width variant of the same architecture (fast parity tests)."""
    p = copy.deepcopy(default_params())
    p["unet_config"]["params"].update(model_channels=64, num_head_channels=16, context_dim=64)
    p["control_stage_config"]["params"].update(model_channels=64, num_head_channels=16, control_model_ratio=0.5,
                                               hint_channels=32, context_dim=64)
    p["first_stage_config"]["params"]["ddconfig"]["ch"] = 32
    p["preprocess_config"]["params"].update(in_nc=128, N=48, M=32, slice_num=3, slice_ch=[8, 8, 16], codebook_size=512)
    return p
