"""rdeic_b200 — B200-native (sm_100a) decoder for RDEIC's relay residual diffusion.

Drop-in for the reference's decode path (ShreyasBhaktharam/RDEIC): `RDEIC` model facade,
`SpacedSampler` / `DDIMSampler`, the checkerboard / quantise / index / VQ entropy front end.
All arithmetic runs in hand-written CUDA kernels (rdeic_b200/csrc -> librdeic_b200.so, C ABI in
include/rdeic_b200.h); there is no CPU or PyTorch-eager fallback.
"""
from .model import RDEIC, load_yaml_config  # noqa: F401
from .spaced_sampler_relay import SpacedSampler  # noqa: F401
from .ddim_sampler_relay import DDIMSampler  # noqa: F401

__all__ = ["RDEIC", "SpacedSampler", "DDIMSampler", "load_yaml_config"]
