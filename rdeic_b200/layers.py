"""Residual blocks of the learned compressor on the sm_100a kernels (drop-in for the reference's
model/layers/res_blk.py and model/layers/conv.py).

Every block consumes the reference's state_dict entries unchanged and runs on NHWC bf16
activations through the tcgen05 implicit-GEMM kernel (`ops.conv_gemm`): bias, LeakyReLU / GELU and
the residual add are fused into the GEMM epilogue, 5x5 convolutions are 25 shifted TMA box loads,
the sub-pixel convolutions' PixelShuffle is a weight-row permutation at load plus one copy kernel,
and stride-2 convolutions (3x3 and the 1x1 shortcut) read the input through a tensor map with element
strides 2 — no im2col gather.

Channel counts that are not multiples of 8 (the 5/3- and 4/3-width hidden layers of
EntropyParametersEX, compression_modules.py:91-104) are padded with zero weight rows / columns, so
padded activations are exactly 0 and the arithmetic is unchanged.

Determinism contract: no atomics, fixed tile schedule, fixed k order and a
fixed-order split-K reduction — the same input bits give the same output bits on every call, so
the CDF indexes built while compressing and while decompressing are identical.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops
from .engine import Conv

BF16 = torch.bfloat16
SD = Dict[str, torch.Tensor]
ACT_NONE, ACT_LRELU, ACT_GELU = 0, 3, 4


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def load_conv(sd: SD, name: str, dev, shuffle: bool = False, c1: Optional[int] = None) -> Conv:
    """nn.Conv2d weights -> packed tensor-core operand.  `shuffle`: the conv feeds nn.PixelShuffle(2)
    (conv.py:7-10); its output channel c*4 + 2i + j moves to (2i + j)*C + c so that the shuffle
    becomes a copy of contiguous C-vectors.  `c1`: the input is the concat of two tensors with c1 and
    I - c1 channels (read as two K segments, never materialised)."""
    w = sd[name + ".weight"].to(dev, torch.float32)
    b = sd[name + ".bias"].to(dev, torch.float32)
    O, I, kh, kw = w.shape
    if shuffle:
        w = w.view(O // 4, 4, I, kh, kw).transpose(0, 1).reshape(O, I, kh, kw)
        b = b.view(O // 4, 4).t().reshape(O)
    n_pad = pad8(O)
    if n_pad != O:
        w = torch.cat([w, torch.zeros((n_pad - O, I, kh, kw), device=dev)], 0)
        b = torch.cat([b, torch.zeros(n_pad - O, device=dev)], 0)
    return Conv(ops.pack_conv_weight(w.contiguous(), c1=c1), b.contiguous(), n_pad, kh * kw)


def conv(x: torch.Tensor, c: Conv, act: int = ACT_NONE, slope: float = 0.0, resid: Optional[torch.Tensor] = None,
         out: Optional[torch.Tensor] = None, out_f32: bool = False, x2: Optional[torch.Tensor] = None,
         stride2: bool = False) -> torch.Tensor:
    """act(conv(cat(x, x2)) + bias) [+ resid] on NHWC bf16 (x / x2 may be channel slices of wider buffers)."""
    return ops.conv_gemm(x, c.w, c.n_out, c.taps, a2=x2, bias=c.b, act=act, act_param=slope, resid=resid, out=out,
                         out_f32=out_f32, stride2=stride2)


@dataclass
class ResidualBlock:
    """res_blk.py:65-96."""
    conv1: Conv
    conv2: Conv
    adaptor: Optional[Conv]
    slope: float = 0.01

    @staticmethod
    def load(sd: SD, p: str, dev) -> "ResidualBlock":
        ad = load_conv(sd, p + ".adaptor", dev) if (p + ".adaptor.weight") in sd else None
        return ResidualBlock(load_conv(sd, p + ".conv1", dev), load_conv(sd, p + ".conv2", dev), ad)

    def __call__(self, x: torch.Tensor, out_f32: bool = False) -> torch.Tensor:
        identity = conv(x, self.adaptor) if self.adaptor is not None else x
        h = conv(x, self.conv1, ACT_LRELU, self.slope)
        return conv(h, self.conv2, ACT_LRELU, self.slope, resid=identity, out_f32=out_f32)


@dataclass
class ResidualBlockUpsample:
    """res_blk.py:39-63 (upsample = 2)."""
    subpel: Conv
    conv: Conv
    upsample: Conv

    @staticmethod
    def load(sd: SD, p: str, dev) -> "ResidualBlockUpsample":
        return ResidualBlockUpsample(load_conv(sd, p + ".subpel_conv.0", dev, shuffle=True), load_conv(sd, p + ".conv", dev),
                                     load_conv(sd, p + ".upsample.0", dev, shuffle=True))

    def __call__(self, x: torch.Tensor, out_f32: bool = False) -> torch.Tensor:
        # LeakyReLU is pointwise, so it commutes with the shuffle and rides in the GEMM epilogue
        h = ops.pixel_shuffle2(conv(x, self.subpel, ACT_LRELU, 0.01))
        identity = ops.pixel_shuffle2(conv(x, self.upsample))
        return conv(h, self.conv, ACT_LRELU, 0.1, resid=identity, out_f32=out_f32)


@dataclass
class ResidualBlockWithStride:
    """res_blk.py:6-37 (stride = 2)."""
    conv1: Conv
    conv2: Conv
    downsample: Conv
    c_in: int

    @staticmethod
    def load(sd: SD, p: str, dev) -> "ResidualBlockWithStride":
        return ResidualBlockWithStride(load_conv(sd, p + ".conv1", dev), load_conv(sd, p + ".conv2", dev),
                                       load_conv(sd, p + ".downsample", dev), sd[p + ".conv1.weight"].shape[1])

    def __call__(self, x: torch.Tensor, out_f32: bool = False) -> torch.Tensor:
        from .engine import S2_IM2COL
        if S2_IM2COL:                                                       # bring-up fallback
            B, H, W, C = x.shape
            cp = (C + 63) // 64 * 64
            col = ops.im2col_3x3_s2(x).view(B, H // 2, W // 2, 9 * cp)
            h = ops.conv_gemm(col, self.conv1.w, self.conv1.n_out, 1, bias=self.conv1.b, act=ACT_LRELU, act_param=0.01)
            identity = conv(col[..., 4 * cp:4 * cp + C], self.downsample)
            return conv(h, self.conv2, ACT_LRELU, 0.1, resid=identity, out_f32=out_f32)
        h = conv(x, self.conv1, ACT_LRELU, 0.01, stride2=True)
        identity = conv(x, self.downsample, stride2=True)                   # 1x1, stride 2: the even-even samples
        return conv(h, self.conv2, ACT_LRELU, 0.1, resid=identity, out_f32=out_f32)
