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


@torch.no_grad()
def decode_streams(model, stream_paths: Sequence[str], c_crossattn: List[torch.Tensor], steps: int,
                   sizes: Optional[Sequence] = None, batch_size: int = 8, sampler: str = "ddpm",
                   guidance_scale: float = 1.0) -> List[torch.Tensor]:
    """The receiver side of inference.py:57-87 / inference_partition.py:438-520 for a folder of
    bitstreams: decompress every stream (learned compressor on the GPU, byte coders on the host),
    bucket the conditionings by latent size, relay-decode each bucket in batches and crop the
    padding (inference.py:156) when the original (h, w) `sizes` are given.
    `c_crossattn[0]` is the [1,77,1024] embedding of the empty prompt (inference.py:134), repeated per
    batch.  Returns uint8 HWC tensors on the GPU, in input order."""
    from .utils import group_by_padded_size

    conds = [model.apply_condition_decompress(p) for p in stream_paths]
    latent_sizes = [(c.shape[-2] * 8, c.shape[-1] * 8) for c, _ in conds]
    out: List[Optional[torch.Tensor]] = [None] * len(conds)
    ctx = c_crossattn[0]
    for _, members in group_by_padded_size(latent_sizes, batch_size, scale=8):
        c_latent = torch.cat([conds[i][0] for i in members], 0)
        hint = torch.cat([conds[i][1] for i in members], 0)
        n = c_latent.shape[0]
        cond = {"c_latent": [c_latent], "c_crossattn": [ctx.expand(n, -1, -1).contiguous() if ctx.shape[0] == 1 else ctx[:n]],
                "guide_hint": hint}
        imgs = relay_decode(model, cond, steps, sampler=sampler, guidance_scale=guidance_scale)
        k = 0
        for i in members:
            b = conds[i][0].shape[0]
            img = imgs[k:k + b]
            k += b
            if sizes is not None:
                h, w = sizes[i]
                img = img[:, :h, :w]
            out[i] = img[0] if b == 1 else img
    return out
