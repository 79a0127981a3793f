"""Slice orchestration of the entropy-model front end (reference model/compression.py:151-273):
10 channel slices, each an anchor pass then a non-anchor pass around the host rANS coder.

The learned conv stacks that produce the entropy parameters (`hyper_dec`, `entropy_parameters_*`,
`local_context`, `channel_context`, `g_s`) are SURVEY §8(f) "next" items and are passed in as
callables (the reference modules, or any deterministic stand-in); this module owns what §8(a12)
puts on the path: the slice/phase control flow and every ckbd / quantise / index call, fused and
GPU-resident.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from . import ckbd


class SliceCoder:
    def __init__(self, slice_ch: Sequence[int], gaussian_conditional: ckbd.GaussianConditional,
                 entropy_parameters_anchor: Sequence[Callable], entropy_parameters_nonanchor: Sequence[Callable],
                 local_context: Sequence[Callable], channel_context: Sequence[Optional[Callable]]):
        self.slice_ch = list(slice_ch)
        self.gc = gaussian_conditional
        self.ep_a, self.ep_n = entropy_parameters_anchor, entropy_parameters_nonanchor
        self.local_context, self.channel_context = local_context, channel_context

    def _params(self, idx, hyper_params, y_hat_slices, slice_anchor=None):
        """compression.py:167-201 / 233-262: inputs of the two entropy-parameter nets of slice idx."""
        ctx = [] if idx == 0 else [self.channel_context[idx](torch.cat(y_hat_slices, dim=1))]
        if slice_anchor is None:
            p = self.ep_a[idx](torch.cat(ctx + [hyper_params], dim=1) if ctx else hyper_params)
        else:
            p = self.ep_n[idx](torch.cat([self.local_context[idx](slice_anchor)] + ctx + [hyper_params], dim=1))
        scales, means = p.chunk(2, 1)
        return scales.contiguous(), means.contiguous(), ctx

    @torch.no_grad()
    def compress(self, y: torch.Tensor, hyper_params: torch.Tensor):
        """compression.py:161-206 -> (symbols, indexes, y_hat)."""
        symbols: List[int] = []
        indexes: List[int] = []
        y_hat_slices: List[torch.Tensor] = []
        off = 0
        for idx, c in enumerate(self.slice_ch):
            y_slice = y[:, off:off + c].contiguous()
            off += c
            slice_anchor, slice_nonanchor = ckbd.ckbd_split(y_slice)
            sa, ma, _ = self._params(idx, hyper_params, y_hat_slices)
            slice_anchor = ckbd.compress_anchor(self.gc, slice_anchor, sa, ma, symbols, indexes)
            sn, mn, _ = self._params(idx, hyper_params, y_hat_slices, slice_anchor)
            slice_nonanchor = ckbd.compress_nonanchor(self.gc, slice_nonanchor, sn, mn, symbols, indexes)
            y_hat_slices.append(ckbd.ckbd_merge(slice_anchor, slice_nonanchor))
        return symbols, indexes, torch.cat(y_hat_slices, dim=1)

    @torch.no_grad()
    def decompress(self, hyper_params: torch.Tensor, decoder, cdf, cdf_lengths, offsets) -> torch.Tensor:
        """compression.py:233-266 -> y_hat."""
        y_hat_slices: List[torch.Tensor] = []
        for idx in range(len(self.slice_ch)):
            sa, ma, _ = self._params(idx, hyper_params, y_hat_slices)
            slice_anchor = ckbd.decompress_anchor(self.gc, sa, ma, decoder, cdf, cdf_lengths, offsets)
            sn, mn, _ = self._params(idx, hyper_params, y_hat_slices, slice_anchor)
            slice_nonanchor = ckbd.decompress_nonanchor(self.gc, sn, mn, decoder, cdf, cdf_lengths, offsets)
            y_hat_slices.append(ckbd.ckbd_merge(slice_nonanchor, slice_anchor))
        return torch.cat(y_hat_slices, dim=1)
