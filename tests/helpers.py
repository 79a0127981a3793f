"""Shared seeded-input helpers (identical to tests/golden/make_golden.py `inputs`)."""
import numpy as np
import torch


def inputs(B, h, w, hint_c, ctx_dim, n_noise):
    g = lambda s: torch.Generator().manual_seed(s)
    c_latent = torch.randn(B, 4, h, w, generator=g(7))
    hint = torch.randn(B, hint_c, h, w, generator=g(8))
    ctx = torch.randn(B, 77, ctx_dim, generator=g(9))
    gn = g(231)
    return c_latent, hint, ctx, [torch.randn(B, 4, h, w, generator=gn) for _ in range(n_noise)]


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def psnr(a, b, peak: float) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    mse = float(np.mean((a - b) ** 2))
    return 99.0 if mse == 0 else float(10 * np.log10(peak * peak / mse))
