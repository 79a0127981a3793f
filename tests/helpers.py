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


def entropy_golden_inputs():
    """Identical to tests/golden/make_golden.py::entropy_inputs."""
    g = torch.Generator().manual_seed(41)
    y = torch.randn(2, 8, 6, 10, generator=g) * 6
    y.view(-1)[3] = -0.0
    K, D = 512, 256
    cb = (torch.rand(K, D, generator=g) * 2 - 1) / K
    cb[7] = cb[300]
    pick = torch.randint(0, K, (2 * 3 * 5,), generator=g)
    pick[0] = 300
    z = cb[pick] + 1e-5 * torch.randn(pick.numel(), D, generator=g)
    z[0] = cb[300]
    z = z.reshape(2, 3, 5, D).permute(0, 3, 1, 2).contiguous()
    return y, cb, z
