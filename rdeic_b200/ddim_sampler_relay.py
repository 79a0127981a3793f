"""Drop-in `DDIMSampler` for the relay decode (reference: model/ddim_sampler_relay.py).

Keeps the reference's quirks: timesteps are `range(0, used_timesteps, used_timesteps // S) + 1`
(top step 241 for S=5 although x_T is noised at t=299), noise is drawn every step even at eta=0,
and classifier-free guidance re-runs the *controlled* model on the unconditional dict and only
when `unconditional_conditioning is not None and scale != 1` (reference :187-192).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from . import ops


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """ldm/modules/diffusionmodules/util.py:53-67."""
    if ddim_discr_method == "uniform":
        stride = num_ddpm_timesteps // num_ddim_timesteps
        steps = np.arange(0, num_ddpm_timesteps, stride)
    elif ddim_discr_method == "quad":
        steps = (np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    return steps + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """util.py:70-81 (alphacums: fp32 numpy of the model buffer)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.used_timesteps
        self.schedule = schedule
        self.noise_fn: Optional[Callable[[int, torch.Tensor], torch.Tensor]] = None

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        """reference :23-52; the per-step scalars stay on the host (they feed kernel arguments)."""
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose)
        ac = self.model.alphas_cumprod.detach().float().cpu().numpy()
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = make_ddim_sampling_parameters(
            ac, self.ddim_timesteps, ddim_eta, verbose)
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1. - self.ddim_alphas)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0.,
               score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100,
               unconditional_guidance_scale=1., unconditional_conditioning=None, dynamic_threshold=None,
               ucg_schedule=None, **kwargs):
        """reference :54-120 -> (samples, intermediates)."""
        for name, val in (("mask", mask), ("score_corrector", score_corrector), ("dynamic_threshold", dynamic_threshold)):
            if val is not None:
                raise NotImplementedError(f"DDIMSampler.sample: `{name}` is outside the relay decode path")
        if quantize_x0 or noise_dropout > 0.:
            raise NotImplementedError("DDIMSampler.sample: quantize_x0 / noise_dropout are outside the relay decode path")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback, x_T=x_T,
                                  log_every_t=log_every_t, temperature=temperature,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, ucg_schedule=ucg_schedule)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, callback=None, img_callback=None, log_every_t=100,
                      temperature=1., unconditional_guidance_scale=1., unconditional_conditioning=None,
                      ucg_schedule=None, **kwargs):
        """reference :123-179."""
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device, torch.float32)
        timesteps = self.ddim_timesteps
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        total_steps = timesteps.shape[0]
        for i, step in enumerate(np.flip(timesteps)):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if ucg_schedule is not None:
                assert len(ucg_schedule) == total_steps
                unconditional_guidance_scale = ucg_schedule[i]
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, temperature=temperature,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning, step_i=i)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, temperature=1., unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, step_i: int = 0, **kwargs):
        """reference :181-231."""
        if self.model.parameterization != "eps":
            raise NotImplementedError("only eps-parameterisation is on the RDEIC decode path (rdeic.yaml)")
        e_u = None
        e_t = self.model.apply_model(x, t, c)
        if not (unconditional_conditioning is None or unconditional_guidance_scale == 1.):
            e_u = self.model.apply_model(x, t, unconditional_conditioning)
        f32 = lambda v: float(np.float32(v))
        a_t, a_prev = np.float32(self.ddim_alphas[index]), np.float32(self.ddim_alphas_prev[index])
        sigma = np.float32(self.ddim_sigmas[index])
        # (1 - a_prev - sigma**2).sqrt() evaluated in fp32 like the reference's torch.full tensors
        dir_coef = np.sqrt(np.float32(np.float32(np.float32(1.) - a_prev) - np.float32(sigma * sigma)))
        noise = torch.randn_like(x) if self.noise_fn is None else self.noise_fn(step_i, x).to(x.device, torch.float32)
        return ops.ddim_update(x, e_t, noise, f32(self.ddim_sqrt_one_minus_alphas[index]), f32(np.sqrt(a_t)),
                               f32(np.sqrt(a_prev)), f32(dir_coef), f32(sigma * np.float32(temperature)), eps_uncond=e_u,
                               guidance_scale=float(unconditional_guidance_scale))
