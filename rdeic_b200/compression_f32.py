"""fp32 kernel mode of the learned compressor's conv stacks (reference model/compression.py:22-46,
model/compression_modules.py:7-104, model/layers/res_blk.py:6-96, model/layers/conv.py:7-14).

Why it exists: the decoder must rebuild the encoder's CDF indexes exactly or the arithmetic coder
desynchronises (compression.py:215-273).  The 64 scale bins are ~12 % wide; the bf16 tensor-core nets
move ~3.6 % of the indexes into a neighbouring bin relative to the reference's fp32 nets, so a
reference-encoded stream would not decode.  Everything that feeds `build_indexes` / the means —
`hyper_dec`, `channel_context`, `local_context`, `entropy_parameters_{anchor,nonanchor}` — therefore
runs here: plain fp32 NHWC tensors, CUDA-core fp32 FMA with fp64 folding of the reduction
(csrc/fp32_mode.cu `rdeic_conv_f32`: 1x1 / 3x3 / 5x5, stride 2, LeakyReLU / exact GELU epilogues,
channel-window operands so no torch.cat is materialised).  The analysis / synthesis transforms
(`encoder`, `hyper_enc`, `decoder`, `out`) exist here too for the all-fp32 verification mode
(`Compression(precision="fp32")`); in the default mixed mode they stay on the tcgen05 bf16 kernels
because nothing they produce reaches the coder's tables.

The 5/3- and 4/3-width hidden layers (26, 21, 53, 42, ... channels) are padded to multiples of 4 with
zero weight rows / columns (padded activations are exactly 0, GELU(0) = 0).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib, ops
from ._lib import ConvF32Params

SD = Dict[str, torch.Tensor]
F32 = torch.float32
ACT_NONE, ACT_LRELU, ACT_GELU = 0, 3, 4


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _pix_stride(t: torch.Tensor) -> int:
    """Elements between two pixels of an NHWC tensor (dense, or a channel window of a wider one).  The stride of
    a size-1 dimension is arbitrary in PyTorch (a [B,C,1,1] -> [B,1,1,C] permute reports 1), so take it from the
    innermost pixel dimension that actually has more than one entry."""
    N, H, W, C = t.shape
    if W > 1:
        return t.stride(2)
    if H > 1:
        return t.stride(1)
    if N > 1:
        return t.stride(0)
    return C


_ws: Dict[str, torch.Tensor] = {}


def _workspace(dev) -> torch.Tensor:
    """Caller-owned split-K scratch of the fp32 conv (the library never allocates): 32 MB per device."""
    ws = _ws.get(str(dev))
    if ws is None:
        ws = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
        _ws[str(dev)] = ws
    return ws


class CompressionNetsF32:
    def __init__(self, sd: SD, prefix: str, device):
        self.sd, self.P, self.dev = sd, prefix, torch.device(device)
        self._w: Dict[str, tuple] = {}

    # ---- parameters ------------------------------------------------------------------------------
    def _weight(self, key: str, shuffle: bool, c1: Optional[int]):
        """nn.Conv2d weight/bias -> ([O4][kh][kw][I4] fp32, bias [O4], O4, ksize).  `shuffle`: the conv
        feeds nn.PixelShuffle(2) (conv.py:7-10): output channel c*4 + 2i + j moves to (2i + j)*C + c so the
        shuffle is a copy of contiguous C-vectors.  Layout change and zero padding only."""
        ent = self._w.get(key)
        if ent is None:
            w = self.sd[key + ".weight"].to(self.dev, F32)
            b = self.sd[key + ".bias"].to(self.dev, F32)
            O, I, kh, kw = w.shape
            if shuffle:
                w = w.view(O // 4, 4, I, kh, kw).transpose(0, 1).reshape(O, I, kh, kw)
                b = b.view(O // 4, 4).t().reshape(O)
            O4 = _pad4(O)
            if O4 != O:
                w = torch.cat([w, torch.zeros((O4 - O, I, kh, kw), device=self.dev)], 0)
                b = torch.cat([b, torch.zeros(O4 - O, device=self.dev)], 0)
            if c1 is None and _pad4(I) != I:                  # input = a padded hidden activation
                w = torch.cat([w, torch.zeros((O4, _pad4(I) - I, kh, kw), device=self.dev)], 1)
            ent = (w.permute(0, 2, 3, 1).contiguous(), b.contiguous(), O4, kh)
            self._w[key] = ent
        return ent

    # ---- kernels -----------------------------------------------------------------------------------
    def conv(self, x, key: str, *, x2=None, stride: int = 1, act: int = ACT_NONE, slope: float = 0.0, resid=None,
             out=None, shuffle: bool = False):
        """act(conv(cat(x, x2)) + bias) [+ resid] on NHWC fp32; x / x2 / out may be channel windows of wider
        NHWC buffers."""
        if not x.is_cuda or x.dtype != F32:
            raise _lib.RdeicLibraryError("CompressionNetsF32.conv needs CUDA fp32 tensors; there is no CPU path")
        B, H, W, C1 = x.shape
        C2 = 0 if x2 is None else x2.shape[-1]
        w, b, n_out, k = self._weight(self.P + key, shuffle, C1 if x2 is not None else None)
        if w.shape[3] != C1 + C2:
            raise ValueError(f"{key}: weight expects {w.shape[3]} input channels, got {C1} + {C2}")
        OH, OW = H // stride, W // stride
        if out is None:
            out = torch.empty((B, OH, OW, n_out), dtype=F32, device=self.dev)
        q = ConvF32Params()
        q.a, q.a_n, q.a_h, q.a_w, q.c1, q.a_ld = _p(x), B, H, W, C1, _pix_stride(x)
        if x2 is not None:
            q.a2, q.c2, q.a2_ld = _p(x2), C2, _pix_stride(x2)
        q.ksize, q.stride, q.up = k, stride, 0
        q.w, q.n_out, q.bias = _p(w), n_out, _p(b)
        if resid is not None:
            q.resid, q.ld_resid = _p(resid), _pix_stride(resid)
        q.alpha, q.act, q.act_param = 1.0, act, slope
        q.out, q.ldo = _p(out), _pix_stride(out)
        ws = _workspace(self.dev)
        q.workspace, q.workspace_bytes = _p(ws), ws.numel()
        ops.check(_lib.load().rdeic_conv_f32(C.byref(q), torch.cuda.current_stream().cuda_stream), "rdeic_conv_f32")
        return out

    @staticmethod
    def _shuffle2(x):
        """nn.PixelShuffle(2) on NHWC [B,H,W,4C] whose channels are ordered (i, j, c): layout change only."""
        B, H, W, C4 = x.shape
        c = C4 // 4
        return x.view(B, H, W, 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, c)

    # ---- blocks (res_blk.py) -------------------------------------------------------------------------
    def residual_block(self, p: str, x, slope: float = 0.01):
        """res_blk.py:65-96."""
        identity = self.conv(x, p + ".adaptor") if (self.P + p + ".adaptor.weight") in self.sd else x
        h = self.conv(x, p + ".conv1", act=ACT_LRELU, slope=slope)
        return self.conv(h, p + ".conv2", act=ACT_LRELU, slope=slope, resid=identity)

    def residual_block_upsample(self, p: str, x):
        """res_blk.py:39-63; LeakyReLU is pointwise, so it commutes with the shuffle and rides in the epilogue."""
        h = self._shuffle2(self.conv(x, p + ".subpel_conv.0", act=ACT_LRELU, slope=0.01, shuffle=True))
        identity = self._shuffle2(self.conv(x, p + ".upsample.0", shuffle=True))
        return self.conv(h, p + ".conv", act=ACT_LRELU, slope=0.1, resid=identity)

    def residual_block_with_stride(self, p: str, x):
        """res_blk.py:6-37."""
        h = self.conv(x, p + ".conv1", stride=2, act=ACT_LRELU, slope=0.01)
        identity = self.conv(x, p + ".downsample", stride=2)
        return self.conv(h, p + ".conv2", act=ACT_LRELU, slope=0.1, resid=identity)

    def _mlp3(self, p: str, x, x2=None, out=None):
        """compression_modules.py:75-104: conv, GELU, conv, GELU, conv at Sequential indices 0, 2, 4."""
        h = self.conv(x, p + "0", x2=x2, act=ACT_GELU)
        h = self.conv(h, p + "2", act=ACT_GELU)
        return self.conv(h, p + "4", out=out)

    # ---- stacks (compression_modules.py, compression.py:22-46) -----------------------------------------------
    def hyper_decoder(self, z_q):
        """compression_modules.py:60-72: z_q [B,h/4,w/4,N] -> hyper_params [B,h,w,2M]."""
        h = "hyper_dec.hyper_dec."
        x = self.residual_block_upsample(h + "0", z_q)
        x = self.residual_block_upsample(h + "1", x)
        x = self.residual_block(h + "2", x)
        return self.residual_block(h + "3", x)

    def channel_context(self, idx: int, y_hat_prefix, out=None):
        return self._mlp3(f"channel_context.{idx}.fushion.", y_hat_prefix, out=out)

    def local_context(self, idx: int, anchor, out=None):
        return self.conv(anchor, f"local_context.{idx}", out=out)

    def entropy_parameters(self, which: str, idx: int, x, x2=None):
        return self._mlp3(f"entropy_parameters_{which}.{idx}.fusion.", x, x2=x2)

    def decoder(self, y_hat):
        """compression_modules.py:27-43 g_s -> guide_hint."""
        g = "decoder.g_s."
        x = self.conv(y_hat, g + "0")
        for i in (1, 2, 3):
            x = self.residual_block(g + str(i), x)
        x = self.residual_block_upsample(g + "4", x)
        for i in (5, 6, 7, 8):
            x = self.residual_block(g + str(i), x)
        return x

    def out(self, guide_hint):
        return self.conv(guide_hint, "out")

    def encoder(self, x):
        """compression_modules.py:7-24 g_a."""
        g = "encoder.g_a."
        for i in (0, 1, 2, 3):
            x = self.residual_block(g + str(i), x)
        x = self.residual_block_with_stride(g + "4", x)
        for i in (5, 6, 7):
            x = self.residual_block(g + str(i), x)
        return self.conv(x, g + "8")

    def hyper_encoder(self, y):
        """compression_modules.py:46-58."""
        h = "hyper_enc.hyper_enc."
        y = self.residual_block(h + "0", y)
        y = self.residual_block(h + "1", y)
        y = self.residual_block_with_stride(h + "2", y)
        return self.residual_block_with_stride(h + "3", y)
