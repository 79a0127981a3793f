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
        self._params_sync(y_hat_slices)
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

    def decompress_with(self, handoff: Callable[[int, torch.Tensor], torch.Tensor]) -> torch.Tensor:
        """compression.py:233-266 with the coder call abstracted: `handoff(k, indexes) -> symbols` (both
        int32 device tensors) is phase k's trip to the host coder.  Everything between two hand-offs is
        a fixed sequence of launches on the current stream (one CUDA-graph segment).  Returns y_hat as
        the NHWC bf16 buffer the synthesis transform reads."""
        gc = self.gc
        y_hat_slices: List[torch.Tensor] = []
        k = 0
        for idx in range(len(self.slice_ch)):
            sa, ma, _ = self._params(idx, None, y_hat_slices)
            msq, ind = ops.ckbd_squeeze_indexes(sa, ma, gc.scale_table, gc.scale_bound, ckbd.ANCHOR)
            slice_anchor = ops.ckbd_decode_phase(handoff(k, ind).view(msq.shape), msq, ckbd.ANCHOR)
            sn, mn, _ = self._params(idx, None, y_hat_slices, slice_anchor)
            msq, ind = ops.ckbd_squeeze_indexes(sn, mn, gc.scale_table, gc.scale_bound, ckbd.NONANCHOR)
            slice_nonanchor = ops.ckbd_decode_phase(handoff(k + 1, ind).view(msq.shape), msq, ckbd.NONANCHOR)
            k += 2
            y_hat_slices.append(ckbd.ckbd_merge(slice_nonanchor, slice_anchor))
        self._params_sync(y_hat_slices)
        return self.y_hat

    def _params_sync(self, y_hat_slices):
        while self.synced < len(y_hat_slices):             # slices finished since the last call
            ops.nchw_to_nhwc_bf16(y_hat_slices[self.synced], self.y_hat, c_off=sum(self.slice_ch[:self.synced]))
            self.synced += 1


class _Plan:
    """CUDA-graph plan of one (batch, z-shape): the launch sequences between host hand-offs are
    captured once and replayed (≈250 launches per image collapse into 21 graph launches for
    decompress, 1 for compress).  Hand-off buffers are per-phase pinned int32 tensors whose copies are
    memcpy nodes of the graphs."""

    def __init__(self, owner: "Compression"):
        self.owner = owner
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.pool = torch.cuda.graph_pool_handle()
        self._ctx = None

    def _begin(self):
        g = torch.cuda.CUDAGraph()
        self.graphs.append(g)
        self._ctx = torch.cuda.graph(g, pool=self.pool)
        self._ctx.__enter__()

    def _end(self, *exc):
        ctx, self._ctx = self._ctx, None
        ctx.__exit__(*(exc or (None, None, None)))


class _DecompressPlan(_Plan):
    def __init__(self, owner: "Compression", B: int, hz: int, wz: int):
        super().__init__(owner)
        dev = owner.device
        self.z_q = torch.zeros((B, owner.N, hz, wz), dtype=torch.float32, device=dev)
        h, w = 4 * hz, 4 * wz
        sizes = [B * c * h * (w // 2) for c in owner.slice_ch for _ in (0, 1)]
        self.idx_host = [torch.empty(n, dtype=torch.int32).pin_memory() for n in sizes]
        self.sym_host = [torch.zeros(n, dtype=torch.int32).pin_memory() for n in sizes]
        owner._decompress_body(self.z_q, self._eager_zero)          # warm-up: lazy workspaces, module loading
        torch.cuda.synchronize()
        self._begin()
        try:
            self.c_latent, self.guide_hint = owner._decompress_body(self.z_q, self._capture_handoff)
        except BaseException as e:
            self._end(type(e), e, e.__traceback__)
            raise
        self._end()

    def _eager_zero(self, k, ind):
        return torch.zeros_like(ind)

    def _capture_handoff(self, k, ind):
        self.idx_host[k].copy_(ind.reshape(-1), non_blocking=True)
        self._end()                                            # graph k ends with the index D2H ...
        self._begin()                                          # ... graph k+1 starts with the symbol H2D
        sym = torch.empty(ind.shape, dtype=torch.int32, device=ind.device)
        sym.copy_(self.sym_host[k].view(ind.shape), non_blocking=True)
        return sym

    def run(self, z_q: torch.Tensor, decode: Callable):
        self.z_q.copy_(z_q)
        stream = torch.cuda.current_stream()
        for k in range(len(self.idx_host)):
            self.graphs[k].replay()
            stream.synchronize()
            self.sym_host[k].numpy()[:] = decode(self.idx_host[k])
        self.graphs[-1].replay()
        return self.c_latent.clone(), self.guide_hint.clone()


class _CompressPlan(_Plan):
    def __init__(self, owner: "Compression", B: int, hx: int, wx: int):
        super().__init__(owner)
        dev = owner.device
        self.x = torch.zeros((B, owner.in_nc, hx, wx), dtype=torch.float32, device=dev)
        n = B * owner.M * (hx // 2) * (wx // 2)
        self.sym_host = torch.empty(n, dtype=torch.int32).pin_memory()
        self.idx_host = torch.empty(n, dtype=torch.int32).pin_memory()
        self.z_host = torch.empty((B, hx // 8, wx // 8), dtype=torch.int64).pin_memory()
        owner._compress_body(self.x)
        torch.cuda.synchronize()
        self._begin()
        try:
            sym, idx, z_idx = owner._compress_body(self.x)
            self.sym_host.copy_(sym.buf, non_blocking=True)
            self.idx_host.copy_(idx.buf, non_blocking=True)
            self.z_host.copy_(z_idx, non_blocking=True)
        except BaseException as e:
            self._end(type(e), e, e.__traceback__)
            raise
        self._end()

    def run(self, x: torch.Tensor):
        self.x.copy_(x)
        self.graphs[0].replay()
        torch.cuda.current_stream().synchronize()
        return self.sym_host.numpy(), self.idx_host.numpy(), self.z_host


class Compression:
    """model/compression.py:10-284.  Tensors cross this API as the reference's NCHW fp32."""

    def __init__(self, in_nc, out_nc, N, M, slice_num, slice_ch, codebook_size, device="cuda",
                 rans_encoder: Optional[Callable] = None, rans_decoder: Optional[Callable] = None,
                 hyper_latent_coder=None, use_cuda_graph: bool = True):
        assert slice_num == len(slice_ch) and sum(slice_ch) == M, "slice_ch must partition the M latent channels"
        self.in_nc, self.out_nc, self.N, self.M = in_nc, out_nc, N, M
        self.slice_num, self.slice_ch = slice_num, list(slice_ch)
        self.codebook_size = codebook_size
        self.device = torch.device(device)
        self.quantize = VectorQuantiser(codebook_size, N, device=device)
        self.gaussian_conditional = ckbd.GaussianConditional(device=device)
        self._rans_encoder, self._rans_decoder, self._hyper_coder = rans_encoder, rans_decoder, hyper_latent_coder
        self.encoder = self.hyper_enc = self.hyper_dec = self.decoder = self.out = None
        self.use_cuda_graph = use_cuda_graph
        self._plans: Dict = {}

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
        self._plans.clear()
        return self

    def update(self, scale_table=None, force=False):
        """compression.py:275-280."""
        self._plans.clear()                                  # the table is baked into captured launches
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
        return self._synthesis_nhwc(ops.nchw_to_nhwc_bf16(y_hat))

    def _synthesis_nhwc(self, y_hat_nhwc: torch.Tensor):
        gh = self.decoder(y_hat_nhwc, out_f32=True)                                          # compression.py:268
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
    def _compress_body(self, x: torch.Tensor):
        """Everything of compression.py:152-203 that runs on the GPU: no host synchronisation inside."""
        y, z = self.analysis(x)
        z_q, encoding_indices = self.quantize.quant(z)
        coder = _FusedSliceCoder(self, self._hyper_params(z_q))
        n_sym = y.numel()                                    # one symbol per latent element
        symbols, indexes = ckbd.SymbolStream(n_sym, self.device), ckbd.SymbolStream(n_sym, self.device)
        coder.compress(y, None, symbols, indexes)
        return symbols, indexes, encoding_indices

    def _decompress_body(self, z_q: torch.Tensor, handoff: Callable):
        coder = _FusedSliceCoder(self, self._hyper_params(z_q))
        return self._synthesis_nhwc(coder.decompress_with(handoff))

    @torch.no_grad()
    def compress(self, x: torch.Tensor):
        """compression.py:151-213."""
        enc_cls, _, hyp = self._coders()
        x = x.to(self.device, torch.float32)
        if x.dim() != 4 or x.shape[1] != self.in_nc or x.shape[2] % 8 or x.shape[3] % 8:
            raise ValueError(f"Compression.compress: expected [B,{self.in_nc},h,w] with h, w multiples of 8, got {tuple(x.shape)}")
        if self.use_cuda_graph:
            key = ("c",) + tuple(x.shape)
            if key not in self._plans:
                self._plans[key] = _CompressPlan(self, x.shape[0], x.shape[2], x.shape[3])
            symbols, indexes, encoding_indices = self._plans[key].run(x)
        else:
            sym, idx, encoding_indices = self._compress_body(x)
            symbols, indexes = sym.host("sym_all"), idx.host("idx_all")
        z_strings = hyp.compress(encoding_indices)
        encoder = enc_cls()
        if not getattr(encoder, "accepts_arrays", False):    # compressai's pybind coder takes lists
            symbols, indexes = symbols.tolist(), indexes.tolist()
        encoder.encode_with_indexes(symbols, indexes, *self._cdfs())
        return {"strings": [[encoder.flush()], [z_strings]], "shape": encoding_indices.shape[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        """compression.py:215-273 -> (c_latent [B,out_nc,2h,2w], guide_hint [B,M,2h,2w])."""
        _, dec_cls, hyp = self._coders()
        y_strings, z_strings = strings[0][0], strings[1][0]
        encoding_indices = hyp.decompress(z_strings, shape)
        z_q = self.quantize.get_codebook_entry(encoding_indices.long())
        decoder = dec_cls()
        decoder.set_stream(y_strings)
        cdfs = self._cdfs()
        arrays = getattr(decoder, "accepts_arrays", False)

        def decode(idx_host: torch.Tensor):
            """pinned int32 indexes -> the coder's symbols (int32 array-like)"""
            return decoder.decode_stream(idx_host.numpy() if arrays else idx_host.tolist(), *cdfs)

        if self.use_cuda_graph:
            key = ("d",) + tuple(z_q.shape)
            if key not in self._plans:
                self._plans[key] = _DecompressPlan(self, z_q.shape[0], z_q.shape[2], z_q.shape[3])
            return self._plans[key].run(z_q, decode)

        def handoff(k, ind):
            hi = ckbd._pinned.to_host(ind, "idx")
            ckbd._pinned.sync()
            hs = ckbd._pinned.get(ind.numel(), "sym_in")
            hs.numpy()[:] = decode(hi)
            return hs.to(ind.device, non_blocking=True).view(ind.shape)

        return self._decompress_body(z_q, handoff)
