"""Slice orchestration of the entropy-model front end (reference model/compression.py:151-273):
10 channel slices, each an anchor pass then a non-anchor pass around the host rANS coder.

`SliceCoder` owns what §8(a12) puts on the path: the slice/phase control flow and every ckbd /
quantise / index call, fused and GPU-resident; the learned conv stacks are callables there (the
reference modules, or any deterministic stand-in).

`Compression` is the drop-in for the reference class of the same name (§8f ranks 1-3): the same
constructor arguments, state_dict keys, `compress(x)` / `decompress(strings, shape)` / `update()`
entry points, with every conv stack (`encoder`, `hyper_enc`, `hyper_dec`, `entropy_parameters_*`,
`local_context`, `channel_context`, `decoder`, `out`) on the tcgen05 implicit-GEMM kernel.  The byte
coders stay in the host libraries the reference uses (compressai rANS, torchac): they are
constructor arguments, resolved from those packages by default.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import ckbd, ops
from .compression_modules import (ChannelContextEX, Decoder, Encoder, EntropyParametersEX, HyperDecoder, HyperEncoder,
                                  VectorQuantiser)
from .layers import conv, load_conv

BF16 = torch.bfloat16


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
    def compress(self, y: torch.Tensor, hyper_params: torch.Tensor, symbols=None, indexes=None):
        """compression.py:161-206 -> (symbols, indexes, y_hat).  `symbols` / `indexes` default to the
        reference's Python lists; pass two `ckbd.SymbolStream`s to keep them on the GPU."""
        symbols = [] if symbols is None else symbols
        indexes = [] if indexes is None else indexes
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


class _FusedSliceCoder(SliceCoder):
    """SliceCoder whose entropy-parameter inputs never leave NHWC bf16: `y_hat` so far lives in one
    [B,h,w,M] buffer (the channel-context nets read its channel prefix in place), each slice's
    local / channel context lands in adjacent channel windows of one buffer, and the torch.cat with
    `hyper_params` (compression.py:170,184,190) is the kernel's second K segment."""

    def __init__(self, owner: "Compression", hyper_nhwc: torch.Tensor):
        super().__init__(owner.slice_ch, owner.gaussian_conditional, owner.entropy_parameters_anchor,
                         owner.entropy_parameters_nonanchor, owner.local_context, owner.channel_context)
        B, h, w, _ = hyper_nhwc.shape
        self.hyper = hyper_nhwc
        self.y_hat = torch.zeros((B, h, w, sum(self.slice_ch)), dtype=BF16, device=hyper_nhwc.device)
        self.ctx = [torch.empty((B, h, w, 4 * c if i else 2 * c), dtype=BF16, device=hyper_nhwc.device)
                    for i, c in enumerate(self.slice_ch)]
        self.synced = 0

    def _params(self, idx, hyper_params, y_hat_slices, slice_anchor=None):
        c = self.slice_ch[idx]
        while self.synced < len(y_hat_slices):             # slices finished since the last call
            ops.nchw_to_nhwc_bf16(y_hat_slices[self.synced], self.y_hat, c_off=sum(self.slice_ch[:self.synced]))
            self.synced += 1
        ctx = self.ctx[idx]
        if slice_anchor is None:
            if idx == 0:
                p = self.ep_a[0](self.hyper, out_f32=True)
            else:
                self.channel_context[idx](self.y_hat[..., :sum(self.slice_ch[:idx])], out=ctx[..., 2 * c:])
                p = self.ep_a[idx](ctx[..., 2 * c:], x2=self.hyper, out_f32=True)
        else:
            conv(ops.nchw_to_nhwc_bf16(slice_anchor), self.local_context[idx], out=ctx[..., :2 * c])
            p = self.ep_n[idx](ctx, x2=self.hyper, out_f32=True)
        return ops.nhwc_to_nchw_f32(p[..., :c]), ops.nhwc_to_nchw_f32(p[..., c:2 * c]), None


class Compression:
    """model/compression.py:10-284.  Tensors cross this API as the reference's NCHW fp32."""

    def __init__(self, in_nc, out_nc, N, M, slice_num, slice_ch, codebook_size, device="cuda",
                 rans_encoder: Optional[Callable] = None, rans_decoder: Optional[Callable] = None,
                 hyper_latent_coder=None):
        assert slice_num == len(slice_ch) and sum(slice_ch) == M, "slice_ch must partition the M latent channels"
        self.in_nc, self.out_nc, self.N, self.M = in_nc, out_nc, N, M
        self.slice_num, self.slice_ch = slice_num, list(slice_ch)
        self.codebook_size = codebook_size
        self.device = torch.device(device)
        self.quantize = VectorQuantiser(codebook_size, N, device=device)
        self.gaussian_conditional = ckbd.GaussianConditional(device=device)
        self._rans_encoder, self._rans_decoder, self._hyper_coder = rans_encoder, rans_decoder, hyper_latent_coder
        self.encoder = self.hyper_enc = self.hyper_dec = self.decoder = self.out = None

    # ---- weights ---------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = "preprocess_model.", strict: bool = True):
        """Takes the RDEIC checkpoint (keys under `preprocess_model.`) or a bare Compression state_dict
        (`prefix=""`).  The analysis side (`encoder`, `hyper_enc`) is optional: a decode-only
        deployment may ship without it."""
        if "state_dict" in sd:
            sd = sd["state_dict"]
        if not any(k.startswith(prefix) for k in sd) and any(k.startswith("module." + prefix) for k in sd):
            prefix = "module." + prefix
        dev, P, sc, M = self.device, prefix, self.slice_ch, self.M
        self.hyper_dec = HyperDecoder(sd, P + "hyper_dec.", dev)
        self.decoder = Decoder(sd, P + "decoder.", dev)
        self.out = load_conv(sd, P + "out", dev)
        self.local_context = [load_conv(sd, f"{P}local_context.{i}", dev) for i in range(self.slice_num)]
        self.channel_context = [ChannelContextEX(sd, f"{P}channel_context.{i}.", dev) if i else None
                                for i in range(self.slice_num)]
        self.entropy_parameters_anchor = [EntropyParametersEX(sd, f"{P}entropy_parameters_anchor.{i}.", dev,
                                                              c1=2 * sc[i] if i else None) for i in range(self.slice_num)]
        self.entropy_parameters_nonanchor = [EntropyParametersEX(sd, f"{P}entropy_parameters_nonanchor.{i}.", dev,
                                                                 c1=4 * sc[i] if i else 2 * sc[i])
                                             for i in range(self.slice_num)]
        self.quantize.load_state_dict(sd, P + "quantize.")
        if (P + "encoder.g_a.0.conv1.weight") in sd:
            self.encoder = Encoder(sd, P + "encoder.", dev)
            self.hyper_enc = HyperEncoder(sd, P + "hyper_enc.", dev)
        elif strict:
            raise KeyError(f"{P}encoder.* is missing from the checkpoint (pass strict=False for decode-only use)")
        return self

    def update(self, scale_table=None, force=False):
        """compression.py:275-280."""
        if scale_table is None:
            scale_table = ckbd.get_scale_table()
        return self.gaussian_conditional.update_scale_table(scale_table, force=force)

    # ---- host coders (out of scope: compressai rANS, torchac) ---------------------------------
    def _coders(self):
        enc, dec, hyp = self._rans_encoder, self._rans_decoder, self._hyper_coder
        if enc is None or dec is None:
            try:
                from compressai.ans import BufferedRansEncoder, RansDecoder
            except ImportError as e:
                raise RuntimeError("Compression needs the host rANS coder: install compressai, or pass "
                                   "rans_encoder= / rans_decoder= factories (compression.py:6)") from e
            enc, dec = enc or BufferedRansEncoder, dec or RansDecoder
        if hyp is None:
            hyp = ckbd.TorchacHyperLatentCoder(self.codebook_size)
        return enc, dec, hyp

    def _cdfs(self):
        return self.gaussian_conditional.cdf_tables()

    # ---- nets ------------------------------------------------------------------------------------
    def _hyper_params(self, z_q: torch.Tensor) -> torch.Tensor:
        return self.hyper_dec(ops.nchw_to_nhwc_bf16(z_q))

    def _synthesis(self, y_hat: torch.Tensor):
        gh = self.decoder(ops.nchw_to_nhwc_bf16(y_hat), out_f32=True)                       # compression.py:268
        gh16 = ops.f32_to_bf16(gh)
        c_latent = conv(gh16, self.out, out_f32=True)                                        # :270
        return ops.nhwc_to_nchw_f32(c_latent, self.out_nc), ops.nhwc_to_nchw_f32(gh)

    @torch.no_grad()
    def analysis(self, x: torch.Tensor):
        """compression.py:152-154 -> (y, z) NCHW fp32."""
        if self.encoder is None:
            raise RuntimeError("this Compression was loaded without the analysis side (encoder.*, hyper_enc.*)")
        y = self.encoder(ops.nchw_to_nhwc_bf16(x.to(self.device, torch.float32)), out_f32=True)
        z = self.hyper_enc(ops.f32_to_bf16(y), out_f32=True)
        return ops.nhwc_to_nchw_f32(y), ops.nhwc_to_nchw_f32(z)

    # ---- reference entry points ----------------------------------------------------------------
    @torch.no_grad()
    def compress(self, x: torch.Tensor):
        """compression.py:151-213."""
        enc_cls, _, hyp = self._coders()
        y, z = self.analysis(x)
        z_q, encoding_indices = self.quantize.quant(z)
        z_strings = hyp.compress(encoding_indices)
        coder = _FusedSliceCoder(self, self._hyper_params(z_q))
        n_sym = y.numel()                                    # one symbol per latent element
        symbols, indexes = ckbd.SymbolStream(n_sym, self.device), ckbd.SymbolStream(n_sym, self.device)
        coder.compress(y, None, symbols, indexes)            # no host synchronisation inside
        symbols, indexes = symbols.host("sym_all"), indexes.host("idx_all")
        encoder = enc_cls()
        if not getattr(encoder, "accepts_arrays", False):    # compressai's pybind coder takes lists
            symbols, indexes = symbols.tolist(), indexes.tolist()
        encoder.encode_with_indexes(symbols, indexes, *self._cdfs())
        return {"strings": [[encoder.flush()], [z_strings]], "shape": z.shape[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        """compression.py:215-273 -> (c_latent [B,out_nc,2h,2w], guide_hint [B,M,2h,2w])."""
        _, dec_cls, hyp = self._coders()
        y_strings, z_strings = strings[0][0], strings[1][0]
        encoding_indices = hyp.decompress(z_strings, shape)
        z_q = self.quantize.get_codebook_entry(encoding_indices.long())
        decoder = dec_cls()
        decoder.set_stream(y_strings)
        coder = _FusedSliceCoder(self, self._hyper_params(z_q))
        y_hat = coder.decompress(None, decoder, *self._cdfs())
        return self._synthesis(y_hat)
