"""Drop-in for the decode-path part of `model/compression_modules.py`: the CVQ-VAE
`VectorQuantiser` look-up (`quant`, `get_codebook_entry`), on CUDA, bit-exact indices."""
from __future__ import annotations

import torch

from . import ops


class VectorQuantiser:
    """model/compression_modules.py:189-338 (inference entry points only).  `load_state_dict` takes
    the reference keys `embedding.weight` [num_embed, embed_dim] and `embed_prob`."""

    def __init__(self, num_embed: int, embed_dim: int, device="cuda", **unused):
        self.num_embed, self.embed_dim = num_embed, embed_dim
        self.device = torch.device(device)
        g = torch.Generator().manual_seed(0)
        w = (torch.rand(num_embed, embed_dim, generator=g) * 2 - 1) / num_embed     # :217 uniform(-1/K, 1/K)
        self.weight = w.to(self.device).contiguous()
        self.embed_prob = torch.zeros(num_embed, device=self.device)

    def load_state_dict(self, sd, prefix: str = ""):
        self.weight = sd[prefix + "embedding.weight"].to(self.device, torch.float32).contiguous()
        if prefix + "embed_prob" in sd:
            self.embed_prob = sd[prefix + "embed_prob"].to(self.device)
        self.num_embed, self.embed_dim = self.weight.shape
        return self

    @torch.no_grad()
    def quant(self, z, temp=None, rescale_logits=False, return_logits=False):
        """:309-331 -> (z_q [B,D,h,w], encoding_indices [B,h,w] int64)."""
        assert temp is None or temp == 1.0, "Only for interface compatible with Gumbel"
        assert rescale_logits is False, "Only for interface compatible with Gumbel"
        assert return_logits is False, "Only for interface compatible with Gumbel"
        return ops.vq_quant(z.to(self.device, torch.float32), self.weight)

    @torch.no_grad()
    def get_codebook_entry(self, indices):
        """:333-338 -> [B,D,h,w]; out-of-range indices raise IndexError like nn.Embedding."""
        indices = indices.to(self.device, torch.int64)
        if indices.numel():
            lo, hi = int(indices.min()), int(indices.max())
            if lo < 0 or hi >= self.num_embed:
                raise IndexError("index out of range in self")
        return ops.vq_lookup(indices, self.weight)
