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
from .compression_f32 import CompressionNetsF32
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


class _F32SliceCoder(_FusedSliceCoder):
    """The same fused data flow with every tensor that reaches `build_indexes` / the means in fp32 on the
    fp32 kernels (`compression_f32.CompressionNetsF32`): `y_hat`, the context windows and `hyper_params`
    are NHWC fp32 buffers, the torch.cat inputs of compression.py:170,184,190 are channel windows / the
    kernel's second source.  This is the coder used by the default (`precision="mixed"`) and the all-fp32
    modes: its CDF indexes agree with the reference's fp32 nets to fp32 round-off."""

    def __init__(self, owner: "Compression", hyper_nhwc: torch.Tensor):
        SliceCoder.__init__(self, owner.slice_ch, owner.gaussian_conditional, None, None, None, None)
        B, h, w, _ = hyper_nhwc.shape
        dev = hyper_nhwc.device
        self.nets = owner.nets32
        self.hyper = hyper_nhwc
        self.y_hat = torch.zeros((B, h, w, sum(self.slice_ch)), dtype=torch.float32, device=dev)
        self.ctx = [torch.empty((B, h, w, 4 * c if i else 2 * c), dtype=torch.float32, device=dev)
                    for i, c in enumerate(self.slice_ch)]
        self.synced = 0

    def _params(self, idx, hyper_params, y_hat_slices, slice_anchor=None):
        c = self.slice_ch[idx]
        self._params_sync(y_hat_slices)
        ctx = self.ctx[idx]
        if slice_anchor is None:
            if idx == 0:
                p = self.nets.entropy_parameters("anchor", 0, self.hyper)
            else:
                self.nets.channel_context(idx, self.y_hat[..., :sum(self.slice_ch[:idx])], out=ctx[..., 2 * c:])
                p = self.nets.entropy_parameters("anchor", idx, ctx[..., 2 * c:], x2=self.hyper)
        else:
            self.nets.local_context(idx, slice_anchor.permute(0, 2, 3, 1).contiguous(), out=ctx[..., :2 * c])
            p = self.nets.entropy_parameters("nonanchor", idx, ctx, x2=self.hyper)
        # NHWC -> the reference's NCHW for the integer kernels: layout change only
        return p[..., :c].permute(0, 3, 1, 2).contiguous(), p[..., c:2 * c].permute(0, 3, 1, 2).contiguous(), None

    def _params_sync(self, y_hat_slices):
        while self.synced < len(y_hat_slices):
            off = sum(self.slice_ch[:self.synced])
            self.y_hat[..., off:off + self.slice_ch[self.synced]].copy_(y_hat_slices[self.synced].permute(0, 2, 3, 1))
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

    def release(self):
        """Drop the graphs (and with them the private pool's activations) and the pinned hand-off buffers."""
        for g in self.graphs:
            g.reset()
        self.graphs.clear()
        for name in ("idx_host", "sym_host", "z_host", "z_q", "x", "c_latent", "guide_hint"):
            if hasattr(self, name):
                setattr(self, name, None)
        self.pool = None


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
                 hyper_latent_coder=None, use_cuda_graph: bool = True, precision: str = "mixed", max_plans: int = 4):
        """`precision`:
          "mixed" (default) — everything that feeds the coder's CDF indexes and means (`hyper_dec`,
                    `channel_context`, `local_context`, `entropy_parameters_*`) runs on the fp32 kernels, so
                    streams are exchangeable with the reference's fp32 nets; the analysis / synthesis
                    transforms run on the tcgen05 bf16 kernels;
          "bf16"  — every conv stack on the tensor cores: fastest, streams decodable only by this mode;
          "fp32"  — every conv stack on the fp32 kernels (verification mode).
        `max_plans`: CUDA-graph plans kept (LRU) — one per distinct (direction, batch, shape)."""
        assert slice_num == len(slice_ch) and sum(slice_ch) == M, "slice_ch must partition the M latent channels"
        if precision not in ("mixed", "bf16", "fp32"):
            raise ValueError(f"Compression: precision must be 'mixed', 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        self.max_plans = int(max_plans)
        self.nets32 = None
        self.in_nc, self.out_nc, self.N, self.M = in_nc, out_nc, N, M
        self.slice_num, self.slice_ch = slice_num, list(slice_ch)
        self.codebook_size = codebook_size
        self.device = torch.device(device)
        self.quantize = VectorQuantiser(codebook_size, N, device=device)
        self.gaussian_conditional = ckbd.GaussianConditional(device=device)
        self._rans_encoder, self._rans_decoder, self._hyper_coder = rans_encoder, rans_decoder, hyper_latent_coder
        self.encoder = self.hyper_enc = self.hyper_dec = self.decoder = self.out = None
        self.use_cuda_graph = use_cuda_graph
        self._plans: Dict = {}           # insertion-ordered: least recently used first
        self._seen: Dict = {}
        self.has_encoder = False

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
        bf16_transforms = self.precision in ("mixed", "bf16")
        has_encoder = (P + "encoder.g_a.0.conv1.weight") in sd
        if self.precision in ("mixed", "fp32"):
            self.nets32 = CompressionNetsF32(sd, P, dev)
        if bf16_transforms:
            self.decoder = Decoder(sd, P + "decoder.", dev)
            self.out = load_conv(sd, P + "out", dev)
            if has_encoder:
                self.encoder = Encoder(sd, P + "encoder.", dev)
                self.hyper_enc = HyperEncoder(sd, P + "hyper_enc.", dev)
        if self.precision == "bf16":
            self.hyper_dec = HyperDecoder(sd, P + "hyper_dec.", dev)
            self.local_context = [load_conv(sd, f"{P}local_context.{i}", dev) for i in range(self.slice_num)]
            self.channel_context = [ChannelContextEX(sd, f"{P}channel_context.{i}.", dev) if i else None
                                    for i in range(self.slice_num)]
            self.entropy_parameters_anchor = [EntropyParametersEX(sd, f"{P}entropy_parameters_anchor.{i}.", dev,
                                                                  c1=2 * sc[i] if i else None) for i in range(self.slice_num)]
            self.entropy_parameters_nonanchor = [EntropyParametersEX(sd, f"{P}entropy_parameters_nonanchor.{i}.", dev,
                                                                     c1=4 * sc[i] if i else 2 * sc[i])
                                                 for i in range(self.slice_num)]
        self.quantize.load_state_dict(sd, P + "quantize.")
        self.has_encoder = has_encoder
        if not has_encoder and strict:
            raise KeyError(f"{P}encoder.* is missing from the checkpoint (pass strict=False for decode-only use)")
        self._drop_plans()
        return self

    def update(self, scale_table=None, force=False):
        """compression.py:275-280."""
        self._drop_plans()                                   # the table is baked into captured launches
        if scale_table is None:
            scale_table = ckbd.get_scale_table()
        return self.gaussian_conditional.update_scale_table(scale_table, force=force)

    # ---- CUDA-graph plan cache -------------------------------------------------------------------------
    def _drop_plans(self):
        for plan in self._plans.values():
            plan.release()
        self._plans.clear()
        self._seen.clear()

    def _plan(self, key, factory):
        """LRU of at most `max_plans` plans.  A shape runs eagerly the first time it is seen and gets a plan
        (eager warm-up + capture + 21 graphs + a private activation pool + pinned hand-off buffers) only when
        it comes back, so a mixed-resolution dataset (inference_partition.py:400-453 buckets) neither pays a
        capture per one-off shape nor accumulates GPU / pinned memory without bound."""
        if not self.use_cuda_graph or self.max_plans <= 0:
            return None
        plan = self._plans.pop(key, None)
        if plan is None:
            n = self._seen.get(key, 0) + 1
            if len(self._seen) > 64:
                self._seen.clear()
            self._seen[key] = n
            if n < 2:
                return None
            while len(self._plans) >= self.max_plans:
                self._plans.pop(next(iter(self._plans))).release()
            plan = factory()
        self._plans[key] = plan                                # most recently used last
        return plan

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
    def _slice_coder(self, z_q: torch.Tensor) -> _FusedSliceCoder:
        """hyper_dec (compression.py:158,222) + the slice coder of the configured precision."""
        if self.precision == "bf16":
            return _FusedSliceCoder(self, self.hyper_dec(ops.nchw_to_nhwc_bf16(z_q)))
        return _F32SliceCoder(self, self.nets32.hyper_decoder(z_q.permute(0, 2, 3, 1).contiguous()))

    def _hyper_params(self, z_q: torch.Tensor) -> torch.Tensor:
        """hyper_params as NHWC (bf16 in the "bf16" mode, else fp32)."""
        return self._slice_coder(z_q).hyper

    def _synthesis(self, y_hat: torch.Tensor):
        if self.precision == "fp32":
            return self._synthesis_nhwc(y_hat.permute(0, 2, 3, 1).contiguous())
        return self._synthesis_nhwc(ops.nchw_to_nhwc_bf16(y_hat))

    def _synthesis_nhwc(self, y_hat_nhwc: torch.Tensor):
        """compression.py:268-270 -> (c_latent, guide_hint) NCHW fp32."""
        if self.precision == "fp32":
            gh = self.nets32.decoder(y_hat_nhwc)
            c_latent = self.nets32.out(gh)
            return (c_latent[..., :self.out_nc].permute(0, 3, 1, 2).contiguous(), gh.permute(0, 3, 1, 2).contiguous())
        if y_hat_nhwc.dtype == torch.float32:                # mixed: y_hat holds integers + fp32 means
            y_hat_nhwc = ops.f32_to_bf16(y_hat_nhwc)
        gh = self.decoder(y_hat_nhwc, out_f32=True)                                          # compression.py:268
        gh16 = ops.f32_to_bf16(gh)
        c_latent = conv(gh16, self.out, out_f32=True)                                        # :270
        return ops.nhwc_to_nchw_f32(c_latent, self.out_nc), ops.nhwc_to_nchw_f32(gh)

    @torch.no_grad()
    def analysis(self, x: torch.Tensor):
        """compression.py:152-154 -> (y, z) NCHW fp32."""
        if not self.has_encoder:
            raise RuntimeError("this Compression was loaded without the analysis side (encoder.*, hyper_enc.*)")
        x = x.to(self.device, torch.float32)
        if self.precision == "fp32":
            y = self.nets32.encoder(x.permute(0, 2, 3, 1).contiguous())
            z = self.nets32.hyper_encoder(y)
            return y.permute(0, 3, 1, 2).contiguous(), z.permute(0, 3, 1, 2).contiguous()
        y = self.encoder(ops.nchw_to_nhwc_bf16(x), out_f32=True)
        z = self.hyper_enc(ops.f32_to_bf16(y), out_f32=True)
        return ops.nhwc_to_nchw_f32(y), ops.nhwc_to_nchw_f32(z)

    # ---- reference entry points ----------------------------------------------------------------
    def _compress_body(self, x: torch.Tensor):
        """Everything of compression.py:152-203 that runs on the GPU: no host synchronisation inside."""
        y, z = self.analysis(x)
        z_q, encoding_indices = self.quantize.quant(z)
        coder = self._slice_coder(z_q)
        n_sym = y.numel()                                    # one symbol per latent element
        symbols, indexes = ckbd.SymbolStream(n_sym, self.device), ckbd.SymbolStream(n_sym, self.device)
        coder.compress(y, None, symbols, indexes)
        return symbols, indexes, encoding_indices

    def _decompress_body(self, z_q: torch.Tensor, handoff: Callable):
        return self._synthesis_nhwc(self._slice_coder(z_q).decompress_with(handoff))

    @torch.no_grad()
    def compress(self, x: torch.Tensor):
        """compression.py:151-213."""
        enc_cls, _, hyp = self._coders()
        x = x.to(self.device, torch.float32)
        if x.dim() != 4 or x.shape[1] != self.in_nc or x.shape[2] % 8 or x.shape[3] % 8:
            raise ValueError(f"Compression.compress: expected [B,{self.in_nc},h,w] with h, w multiples of 8, got {tuple(x.shape)}")
        plan = self._plan(("c",) + tuple(x.shape), lambda: _CompressPlan(self, x.shape[0], x.shape[2], x.shape[3]))
        if plan is not None:
            symbols, indexes, encoding_indices = plan.run(x)
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

        plan = self._plan(("d",) + tuple(z_q.shape), lambda: _DecompressPlan(self, z_q.shape[0], z_q.shape[2], z_q.shape[3]))
        if plan is not None:
            return plan.run(z_q, decode)

        def handoff(k, ind):
            hi = ckbd._pinned.to_host(ind, "idx")
            ckbd._pinned.sync()
            hs = ckbd._pinned.get(ind.numel(), "sym_in")
            hs.numpy()[:] = decode(hi)
            return hs.to(ind.device, non_blocking=True).view(ind.shape)

        return self._decompress_body(z_q, handoff)
