"""The decode half of the reference driver `process()` (inference.py:57-87,
inference_partition.py:280-310): from decompressed conditioning to uint8 images.

    cond = {"c_latent": [c_latent], "c_crossattn": [ctx], "guide_hint": guide_hint}
    imgs = relay_decode(model, cond, steps=5)          # uint8 [B,H,W,3] on the GPU
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from .ddim_sampler_relay import DDIMSampler
from .spaced_sampler_relay import SpacedSampler


@torch.no_grad()
def relay_decode(model, cond: Dict, steps: int, sampler: str = "ddpm", guidance_scale: float = 1.0,
                 start_noise: Optional[torch.Tensor] = None, step_noises: Optional[Sequence[torch.Tensor]] = None,
                 as_uint8: bool = True) -> torch.Tensor:
    """inference.py:63-87.  `start_noise` / `step_noises` replace the reference's device-side
    torch.randn draws (inference.py:65, spaced_sampler_relay.py:378) when reproducibility across
    devices is needed; by default noise is drawn on the GPU exactly where the reference draws it."""
    c_latent = cond["c_latent"][0].to(model.device, torch.float32)
    n, _, h, w = c_latent.shape
    shape = (n, 4, h, w)
    noise = torch.randn(shape, device=model.device, dtype=torch.float32) if start_noise is None else start_noise
    # inference.py:66-67: t = used_timesteps - 1 for every sample (host list: no device round trip)
    x_T = model.q_sample(x_start=c_latent, t=[model.used_timesteps - 1] * n, noise=noise)
    noise_fn: Optional[Callable] = None
    if step_noises is not None:
        noise_fn = lambda i, like: step_noises[i]
    if sampler == "ddpm":
        s = SpacedSampler(model, var_type="fixed_small")
        s.noise_fn = noise_fn
        samples = s.sample(steps, shape, cond, unconditional_guidance_scale=guidance_scale,
                           unconditional_conditioning=None, cond_fn=None, x_T=x_T)
    else:
        s = DDIMSampler(model)
        s.noise_fn = noise_fn
        samples, _ = s.sample(S=steps, batch_size=n, shape=shape[1:], conditioning=cond,
                              unconditional_conditioning=None, unconditional_guidance_scale=guidance_scale,
                              x_T=x_T, eta=0, verbose=False)
    if as_uint8:
        return model.decode_first_stage_u8(samples)
    return model.decode_first_stage(samples)
