"""Drop-in `SpacedSampler` for the relay decode (reference: model/spaced_sampler_relay.py).

Same constructor and `sample(...)` signature; the respaced schedule is host-side fp64 numpy like
the reference's, the per-step posterior update runs in one fused CUDA kernel
(`rdeic_relay_update`, csrc/elementwise.cu) instead of ~10 ATen launches and per-call H2D table
copies (`_extract_into_tensor`, reference :65-77).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from . import ops


def space_timesteps(num_timesteps: int, section_counts):
    """Pick `section_counts` evenly spread steps out of range(num_timesteps)
    (reference :11-61; guided-diffusion respace)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                picked = range(0, num_timesteps, stride)
                if len(picked) == want:
                    return set(picked)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(s) for s in section_counts.split(",")]
    n_sections = len(section_counts)
    base, rem = divmod(num_timesteps, n_sections)
    out = []
    offset = 0
    for i, count in enumerate(section_counts):
        size = base + (1 if i < rem else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            out.append(offset + round(pos))
            pos += stride
        offset += size
    return set(out)


class SpacedSampler:
    def __init__(self, model, schedule: str = "linear", var_type: str = "fixed_small"):
        self.model = model
        self.original_num_steps = model.num_timesteps
        self.used_num_steps = model.used_timesteps
        self.schedule = schedule
        self.var_type = var_type
        # test / reproducibility hook: noise_fn(step_i, like) -> tensor; default is the reference's
        # torch.randn_like on the model device (reference :378)
        self.noise_fn: Optional[Callable[[int, torch.Tensor], torch.Tensor]] = None

    def make_schedule(self, num_steps: int) -> None:
        """Respaced betas over the first `used_timesteps` of the original chain (reference :88-142)."""
        if self.schedule != "linear":
            raise ValueError(f"schedule '{self.schedule}' unknown.")
        ob = (torch.linspace(self.model.linear_start ** 0.5, self.model.linear_end ** 0.5, self.original_num_steps,
                             dtype=torch.float64) ** 2).numpy()
        oac = np.cumprod(1.0 - ob, axis=0)
        used = space_timesteps(self.used_num_steps, str(num_steps))
        betas, last = [], 1.0
        for i, ac in enumerate(oac):
            if i in used:
                betas.append(1 - ac / last)
                last = ac
        assert len(betas) == num_steps
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        self.timesteps = np.array(sorted(used), dtype=np.int32)
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        assert self.alphas_cumprod_prev.shape == (num_steps,)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:])) \
            if num_steps > 1 else np.log(np.maximum(self.posterior_variance, 1e-20))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    def q_sample(self, x_start, t: int, noise=None):
        """reference :144-152 on the respaced tables (t is a respaced index)."""
        if noise is None:
            noise = torch.randn_like(x_start)
        assert noise.shape == x_start.shape
        return ops.q_sample(x_start, noise, float(np.float32(self.sqrt_alphas_cumprod[t])),
                            float(np.float32(self.sqrt_one_minus_alphas_cumprod[t])))

    @torch.no_grad()
    def sample(self, steps, shape, conditioning=None, x_T=None, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, cond_fn=None):
        """reference :172-191."""
        self.make_schedule(num_steps=steps)
        return self.sapced_sampling(conditioning, shape, x_T=x_T,
                                    unconditional_guidance_scale=unconditional_guidance_scale,
                                    unconditional_conditioning=unconditional_conditioning, cond_fn=cond_fn)

    @torch.no_grad()
    def sapced_sampling(self, cond, shape, x_T, unconditional_guidance_scale, unconditional_conditioning, cond_fn):
        """reference :214-240 (the method name keeps the reference's spelling)."""
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device, torch.float32)
        total_steps = len(self.timesteps)
        for i, step in enumerate(np.flip(self.timesteps)):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            img = self.p_sample_spaced(img, cond, ts, index=index, step_i=i,
                                       unconditional_guidance_scale=unconditional_guidance_scale,
                                       unconditional_conditioning=unconditional_conditioning, cond_fn=cond_fn)
        return img

    def predict_noise(self, x, t, c, unconditional_guidance_scale, unconditional_conditioning):
        """reference :277-290; returns (eps_cond, eps_uncond or None) — the guidance combine is
        fused into the update kernel."""
        if self.model.parameterization != "eps":
            raise NotImplementedError("only eps-parameterisation is on the RDEIC decode path (rdeic.yaml)")
        if unconditional_conditioning is None and unconditional_guidance_scale == 1.:
            return self.model.apply_model(x, t, c), None
        return self.model.apply_model(x, t, c), self.model.apply_model_unconditional(x, t, c)

    @torch.no_grad()
    def p_sample_spaced(self, x, c, t, index, unconditional_guidance_scale, unconditional_conditioning, cond_fn,
                        step_i: int = 0):
        """reference :349-384."""
        if cond_fn is not None:
            raise NotImplementedError("classifier guidance (cond_fn) is outside the relay decode path")
        variance = {"fixed_large": np.append(self.posterior_variance[1], self.betas[1:]) if len(self.betas) > 1
                    else self.betas, "fixed_small": self.posterior_variance}[self.var_type]
        e_t, e_u = self.predict_noise(x, t, c, unconditional_guidance_scale, unconditional_conditioning)
        noise = torch.randn_like(x) if self.noise_fn is None else self.noise_fn(step_i, x).to(x.device, torch.float32)
        f32 = lambda v: float(np.float32(v))
        # nonzero_mask * sqrt(var) with var first cast to fp32, as `_extract_into_tensor(...).float()` does
        sigma = 0.0 if index == 0 else float(np.sqrt(np.float32(variance[index])))
        return ops.relay_update(x, e_t, noise, f32(self.sqrt_recip_alphas_cumprod[index]),
                                f32(self.sqrt_recipm1_alphas_cumprod[index]), f32(self.posterior_mean_coef1[index]),
                                f32(self.posterior_mean_coef2[index]), sigma, eps_uncond=e_u,
                                guidance_scale=float(unconditional_guidance_scale))
