"""Drop-in for `model/compression_modules.py`: the CVQ-VAE `VectorQuantiser` look-up (`quant`,
`get_codebook_entry`; bit-exact indices) and the compressor's conv stacks — analysis / synthesis
transforms, hyper encoder / decoder, channel-context and entropy-parameter nets — on the sm_100a
kernels.  The conv stacks take and return NHWC bf16 tensors (`out_f32` selects an fp32 result for
the tensors that feed integer arithmetic: scales, means, y); `rdeic_b200.compression.Compression`
owns the NCHW fp32 boundary of the reference API."""
from __future__ import annotations

from typing import Dict, List

import torch

from . import ops
from .layers import ACT_GELU, ResidualBlock, ResidualBlockUpsample, ResidualBlockWithStride, conv, load_conv

SD = Dict[str, torch.Tensor]


class _Sequential:
    def __init__(self, blocks: List):
        self.blocks = blocks

    def __call__(self, x: torch.Tensor, out_f32: bool = False) -> torch.Tensor:
        for i, b in enumerate(self.blocks):
            last = i == len(self.blocks) - 1
            x = b(x, out_f32=out_f32 and last)
        return x


class _Conv3x3:
    def __init__(self, sd: SD, p: str, dev):
        self.c = load_conv(sd, p, dev)

    def __call__(self, x, out_f32: bool = False):
        return conv(x, self.c, out_f32=out_f32)


class Encoder(_Sequential):
    """compression_modules.py:7-24 analysis transform g_a: [B,H,W,in_nc] -> y [B,H/2,W/2,M]."""

    def __init__(self, sd: SD, p: str, dev):
        g = p + "g_a."
        super().__init__([ResidualBlock.load(sd, g + str(i), dev) for i in (0, 1, 2, 3)] +
                         [ResidualBlockWithStride.load(sd, g + "4", dev)] +
                         [ResidualBlock.load(sd, g + str(i), dev) for i in (5, 6, 7)] + [_Conv3x3(sd, g + "8", dev)])


class Decoder(_Sequential):
    """compression_modules.py:27-43 synthesis transform g_s: y_hat [B,h,w,M] -> guide_hint [B,2h,2w,M]."""

    def __init__(self, sd: SD, p: str, dev):
        g = p + "g_s."
        super().__init__([_Conv3x3(sd, g + "0", dev)] + [ResidualBlock.load(sd, g + str(i), dev) for i in (1, 2, 3)] +
                         [ResidualBlockUpsample.load(sd, g + "4", dev)] +
                         [ResidualBlock.load(sd, g + str(i), dev) for i in (5, 6, 7, 8)])


class HyperEncoder(_Sequential):
    """compression_modules.py:46-58: y -> z [B,h/4,w/4,N]."""

    def __init__(self, sd: SD, p: str, dev):
        h = p + "hyper_enc."
        super().__init__([ResidualBlock.load(sd, h + "0", dev), ResidualBlock.load(sd, h + "1", dev),
                          ResidualBlockWithStride.load(sd, h + "2", dev), ResidualBlockWithStride.load(sd, h + "3", dev)])


class HyperDecoder(_Sequential):
    """compression_modules.py:60-72: z_q [B,h/4,w/4,N] -> hyper_params [B,h,w,2M]."""

    def __init__(self, sd: SD, p: str, dev):
        h = p + "hyper_dec."
        super().__init__([ResidualBlockUpsample.load(sd, h + "0", dev), ResidualBlockUpsample.load(sd, h + "1", dev),
                          ResidualBlock.load(sd, h + "2", dev), ResidualBlock.load(sd, h + "3", dev)])


class _ConvGeluStack:
    """conv, GELU, conv, GELU, conv at Sequential indices 0, 2, 4."""

    def __init__(self, sd: SD, p: str, dev, c1=None):
        self.c = [load_conv(sd, p + "0", dev, c1=c1), load_conv(sd, p + "2", dev), load_conv(sd, p + "4", dev)]

    def __call__(self, x: torch.Tensor, out=None, out_f32: bool = False, x2=None) -> torch.Tensor:
        x = conv(x, self.c[0], ACT_GELU, x2=x2)
        x = conv(x, self.c[1], ACT_GELU)
        return conv(x, self.c[2], out=out, out_f32=out_f32)


class ChannelContextEX(_ConvGeluStack):
    """compression_modules.py:75-88: three 5x5 convs (25-tap implicit GEMM), in_dim -> 224 -> 128 -> out_dim."""

    def __init__(self, sd: SD, p: str, dev):
        super().__init__(sd, p + "fushion.", dev)


class EntropyParametersEX(_ConvGeluStack):
    """compression_modules.py:90-104: 1x1 MLP in_dim -> 5/3 out -> 4/3 out -> out_dim (scales | means).
    `c1` splits in_dim into (context channels, hyper_params channels): the torch.cat of
    compression.py:170,184,190 is read as two K segments."""

    def __init__(self, sd: SD, p: str, dev, c1=None):
        super().__init__(sd, p + "fusion.", dev, c1=c1)


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
