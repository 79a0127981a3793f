"""fp32 kernel mode of the UNet + control step (BASELINE north_star: per-step rel-L2 <= 1e-5 against the
reference's fp32 output; reference: model/rdeic.py:174-235 over openaimodel.py / attention.py).

Same flat state_dict, same call contract as `NoiseEstimatorEngine.forward`, but every tensor is fp32
NHWC and every contraction runs on the CUDA cores with fp64 folding of the reduction
(csrc/fp32_mode.cu).  This is the verification mode of the path: it shows that the drop-in's
structure (block order, concat order, zero-conv injection, head splitting, GroupNorm groups / eps,
timestep handling) reproduces the reference to fp32 round-off, independently of bf16 effects.  The
throughput mode is `engine.NoiseEstimatorEngine`.

    RDEIC.from_config(cfg, precision="fp32")         # or NoiseEstimatorF32(sd, unet_cfg, ctrl_cfg)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib, ops
from ._lib import ConvF32Params
from .engine import find_denominator

SD = Dict[str, torch.Tensor]
F32 = torch.float32


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class NoiseEstimatorF32:
    def __init__(self, sd: SD, unet_cfg: dict, ctrl_cfg: dict, device="cuda"):
        self.dev = torch.device(device)
        self.sd = sd
        self.model_channels = int(unet_cfg["model_channels"])
        self.base_d_head = int(unet_cfg["num_head_channels"])
        self.ctrl_d_head = int(ctrl_cfg["num_head_channels"])
        self.control_scale = float(ctrl_cfg.get("control_scale", 1.0))
        self._w: Dict[str, torch.Tensor] = {}          # packed fp32 weights [n_out][taps][cin]
        self._v: Dict[str, torch.Tensor] = {}          # fp32 vectors (bias, gamma, beta)
        self._ws = None

    # ---- parameters -------------------------------------------------------------------------------
    def _weight(self, key: str) -> torch.Tensor:
        w = self._w.get(key)
        if w is None:
            t = self.sd[key].to(self.dev, F32)
            if t.dim() == 2:
                t = t[:, :, None, None]
            w = t.permute(0, 2, 3, 1).contiguous()          # OIHW -> [O][kh][kw][I]: layout change only
            self._w[key] = w
        return w

    def _vec(self, key: str) -> Optional[torch.Tensor]:
        if key not in self.sd:
            return None
        v = self._v.get(key)
        if v is None:
            v = self.sd[key].to(self.dev, F32).contiguous()
            self._v[key] = v
        return v

    # ---- kernels ------------------------------------------------------------------------------------
    def conv(self, x, p: str, x2=None, stride=1, up=0, act=0, row_bias=None, resid=None, alpha=1.0, out=None):
        """x NHWC fp32 [B,H,W,C1] (+ x2) -> [B,OH,OW,n_out]"""
        w = self._weight(p + ".weight")
        n_out, k = w.shape[0], w.shape[1]
        B, H, W, C1 = x.shape
        C2 = 0 if x2 is None else x2.shape[-1]
        assert w.shape[3] == C1 + C2, (p, tuple(w.shape), C1, C2)
        OH, OW = (H << up) // stride, (W << up) // stride
        if out is None:
            out = torch.empty((B, OH, OW, n_out), dtype=F32, device=self.dev)
        q = ConvF32Params()
        q.a, q.a_n, q.a_h, q.a_w, q.c1 = _p(x), B, H, W, C1
        q.a2, q.c2 = _p(x2), C2
        q.ksize, q.stride, q.up = k, stride, up
        q.w, q.n_out = _p(w), n_out
        q.bias = _p(self._vec(p + ".bias"))
        if row_bias is not None:
            q.row_bias, q.row_bias_ld = _p(row_bias), row_bias.stride(0)
        if resid is not None:
            q.resid, q.ld_resid = _p(resid), resid.stride(-2)
        q.alpha, q.act = alpha, act
        q.out, q.ldo = _p(out), out.stride(-2)
        ops.check(_lib.load().rdeic_conv_f32(C.byref(q), _stream()), "rdeic_conv_f32")
        return out

    def linear(self, x, p: str, **kw):
        K = x.shape[-1]
        M = x.numel() // K
        resid = kw.pop("resid", None)
        if resid is not None:
            resid = resid.reshape(1, 1, M, resid.shape[-1])
        y = self.conv(x.reshape(1, 1, M, K), p, resid=resid, **kw)
        return y.view(*x.shape[:-1], y.shape[-1])

    def gn(self, x, p: str, eps: float, silu: bool, x2=None):
        B, H, W, C1 = x.shape
        C2 = 0 if x2 is None else x2.shape[-1]
        out = torch.empty((B, H, W, C1 + C2), dtype=F32, device=self.dev)
        if self._ws is None or self._ws.numel() < _lib.load().rdeic_groupnorm_workspace_bytes(B, 1, 8):
            self._ws = torch.empty(_lib.load().rdeic_groupnorm_workspace_bytes(B, 1, 8), dtype=torch.uint8, device=self.dev)
        ops.check(_lib.load().rdeic_groupnorm_nhwc_f32(_p(x), C1, _p(x2), C2, _p(self._vec(p + ".weight")),
                                                       _p(self._vec(p + ".bias")), _p(out), B, H * W,
                                                       find_denominator(C1 + C2, 32), eps, 1 if silu else 0,
                                                       _p(self._ws), _stream()), "rdeic_groupnorm_nhwc_f32")
        return out

    def ln(self, x, p: str):
        Cc = x.shape[-1]
        out = torch.empty_like(x)
        ops.check(_lib.load().rdeic_layernorm_f32(_p(x), _p(self._vec(p + ".weight")), _p(self._vec(p + ".bias")), _p(out),
                                                  x.numel() // Cc, Cc, 1e-5, _stream()), "rdeic_layernorm_f32")
        return out

    def attention(self, q, k, v, heads: int):
        B, Nq, inner = q.shape
        d = inner // heads
        out = torch.empty_like(q)
        ops.check(_lib.load().rdeic_attention_f32(_p(q), _p(k), _p(v), _p(out), B, heads, Nq, k.shape[1], d, q.stride(1),
                                                  k.stride(1), v.stride(1), out.stride(1), q.stride(0), k.stride(0),
                                                  v.stride(0), out.stride(0), d ** -0.5, _stream()), "rdeic_attention_f32")
        return out

    # ---- blocks (same order of operations as the reference modules) ------------------------------------
    def resblock(self, p: str, x, x2, emb_silu):
        """openaimodel.py:249-274 (x2: the skip tensor of the decoder concat, openaimodel.py:804)."""
        h = self.gn(x, p + ".in_layers.0", 1e-5, True, x2=x2)
        rb = self.linear(emb_silu, p + ".emb_layers.1")                       # [B, cout]
        h = self.conv(h, p + ".in_layers.2", row_bias=rb)
        h = self.gn(h, p + ".out_layers.0", 1e-5, True)
        if (p + ".skip_connection.weight") in self.sd:
            xs = self.conv(x, p + ".skip_connection", x2=x2)
        else:
            xs = x
        return self.conv(h, p + ".out_layers.3", resid=xs)

    def transformer(self, p: str, x, context, d_head_cfg: int, is_control: bool):
        """attention.py:331-350 (use_linear, depth 1) + :281-285 + :171-203."""
        B, H, W, Cc = x.shape
        d_head = find_denominator(Cc, d_head_cfg) if is_control else d_head_cfg
        heads = Cc // d_head
        t = p + ".transformer_blocks.0"
        h = self.gn(x, p + ".norm", 1e-6, False).view(B, H * W, Cc)
        h = self.linear(h, p + ".proj_in")
        n = self.ln(h, t + ".norm1")
        a = self.attention(self.linear(n, t + ".attn1.to_q"), self.linear(n, t + ".attn1.to_k"),
                           self.linear(n, t + ".attn1.to_v"), heads)
        h = self.linear(a, t + ".attn1.to_out.0", resid=h)
        n = self.ln(h, t + ".norm2")
        a = self.attention(self.linear(n, t + ".attn2.to_q"), self.linear(context, t + ".attn2.to_k"),
                           self.linear(context, t + ".attn2.to_v"), heads)
        h = self.linear(a, t + ".attn2.to_out.0", resid=h)
        n = self.ln(h, t + ".norm3")
        f = self.linear(n, t + ".ff.net.0.proj")
        g = torch.empty((B, H * W, f.shape[-1] // 2), dtype=F32, device=self.dev)
        ops.check(_lib.load().rdeic_geglu_f32(_p(f), _p(g), B * H * W, g.shape[-1], _stream()), "rdeic_geglu_f32")
        h = self.linear(g, t + ".ff.net.2", resid=h)
        return self.linear(h, p + ".proj_out", resid=x.view(B, H * W, Cc)).view(B, H, W, Cc)

    def block(self, p: str, x, x2, emb_silu, context, d_head_cfg: int, is_control: bool, x_in2=None):
        """TimestepEmbedSequential (openaimodel.py:79-88): children .0, .1, .2 recognised by their keys."""
        i = 0
        while True:
            q = f"{p}.{i}"
            if (q + ".in_layers.0.weight") in self.sd:
                x, x2 = self.resblock(q, x, x2, emb_silu), None
            elif (q + ".proj_in.weight") in self.sd:
                x = self.transformer(q, x, context, d_head_cfg, is_control)
            elif (q + ".op.weight") in self.sd:
                x = self.conv(x, q + ".op", stride=2)
            elif (q + ".conv.weight") in self.sd:
                x = self.conv(x, q + ".conv", up=1)
            elif (q + ".weight") in self.sd and self.sd[q + ".weight"].dim() == 4:
                x = self.conv(x, q, x2=x_in2)
            else:
                return x
            i += 1

    def _count(self, prefix: str) -> int:
        idx = {int(k[len(prefix):].split(".")[0]) for k in self.sd if k.startswith(prefix)}
        return max(idx) + 1 if idx else 0

    def _emb(self, P: str, t_emb):
        """time_embed (openaimodel.py:539-543); every consumer applies SiLU first (emb_layers.0)."""
        e = self.linear(t_emb, P + ".time_embed.0", act=1)
        return self.linear(e, P + ".time_embed.2", act=1)

    @torch.no_grad()
    def forward(self, x, t, context, guide_hint, unconditional: bool = False):
        """x [B,4,h,w] fp32 NCHW, t [B] int64, context [B,77,ctx], guide_hint [B,hint,h,w] -> eps [B,4,h,w]."""
        if not x.is_cuda:
            raise _lib.RdeicLibraryError("NoiseEstimatorF32.forward needs CUDA tensors; there is no CPU path")
        B_, C_ = "model.diffusion_model", "control_model.control_model"
        nhwc = lambda a: a.to(self.dev, F32).permute(0, 2, 3, 1).contiguous()       # layout change only
        hb = nhwc(x)
        ctx = context.to(self.dev, F32).contiguous()
        t_emb = torch.empty((x.shape[0], self.model_channels), dtype=F32, device=self.dev)
        ops.check(_lib.load().rdeic_timestep_embedding_f32(_p(t.to(self.dev, torch.int64).contiguous()), _p(t_emb), x.shape[0],
                                                           self.model_channels, 10000.0, _stream()),
                  "rdeic_timestep_embedding_f32")
        eb = self._emb(B_, t_emb)
        n_in, n_out = self._count(B_ + ".input_blocks."), self._count(B_ + ".output_blocks.")
        hs_base = []
        if unconditional:
            for i in range(n_in):
                hb = self.block(f"{B_}.input_blocks.{i}", hb, None, eb, ctx, self.base_d_head, False)
                hs_base.append(hb)
            hb = self.block(f"{B_}.middle_block", hb, None, eb, ctx, self.base_d_head, False)
            for i in range(n_out):
                hb = self.block(f"{B_}.output_blocks.{i}", hb, hs_base.pop(), eb, ctx, self.base_d_head, False)
        else:
            ec = self._emb(C_, t_emb)
            scales = [float(v) * self.control_scale for v in self.sd["control_model.scale_list"].float().cpu()]
            hint = nhwc(guide_hint)
            hc, hs_ctr, si = hb, [], 0
            for i in range(n_in):
                hb = self.block(f"{B_}.input_blocks.{i}", hb, None, eb, ctx, self.base_d_head, False)
                hc = self.block(f"{C_}.input_blocks.{i}", hc, None, ec, ctx, self.ctrl_d_head, True,
                                x_in2=hint if i == 0 else None)
                hb = self.conv(hc, f"control_model.enc_zero_convs_out.{i}.0", resid=hb, alpha=scales[si])
                si += 1
                hs_base.append(hb)
                hs_ctr.append(hc)
            hb = self.block(f"{B_}.middle_block", hb, None, eb, ctx, self.base_d_head, False)
            hc = self.block(f"{C_}.middle_block", hc, None, ec, ctx, self.ctrl_d_head, True)
            hb = self.conv(hc, "control_model.middle_block_out.0", resid=hb, alpha=scales[si])
            si += 1
            for i in range(n_out):
                hb = self.conv(hs_ctr.pop(), f"control_model.dec_zero_convs_out.{i}.0", resid=hb, alpha=scales[si])
                si += 1
                hb = self.block(f"{B_}.output_blocks.{i}", hb, hs_base.pop(), eb, ctx, self.base_d_head, False)
        h = self.gn(hb, B_ + ".out.0", 1e-5, True)
        return self.conv(h, B_ + ".out.2").permute(0, 3, 1, 2).contiguous()


class VAEDecoderF32(NoiseEstimatorF32):
    """fp32 kernel mode of `decode_first_stage` (ddpm.py:835-844 -> autoencoder.py:97-100 -> model.py:653-686)
    on the same fp32 kernels."""

    def __init__(self, sd: SD, scale_factor: float, device="cuda", prefix: str = "first_stage_model"):
        self.dev = torch.device(device)
        self.sd, self.P, self.scale_factor = sd, prefix, float(scale_factor)
        self._w, self._v, self._ws = {}, {}, None

    def _resnet(self, p: str, x):
        """model.py:128-151 (temb None), GroupNorm eps 1e-6."""
        h = self.conv(self.gn(x, p + ".norm1", 1e-6, True), p + ".conv1")
        xs = self.conv(x, p + ".nin_shortcut") if (p + ".nin_shortcut.weight") in self.sd else x
        return self.conv(self.gn(h, p + ".norm2", 1e-6, True), p + ".conv2", resid=xs)

    def _attn(self, p: str, x):
        """model.py:181-205: one head of width C."""
        B, H, W, Cc = x.shape
        hn = self.gn(x, p + ".norm", 1e-6, False).view(B, H * W, Cc)
        a = self.attention(self.linear(hn, p + ".q"), self.linear(hn, p + ".k"), self.linear(hn, p + ".v"), 1)
        return self.linear(a, p + ".proj_out", resid=x.view(B, H * W, Cc)).view(B, H, W, Cc)

    @torch.no_grad()
    def decode(self, z):
        """[B,4,h,w] fp32 NCHW latent -> [B,3,8h,8w] fp32 in [-1,1]."""
        if not z.is_cuda:
            raise _lib.RdeicLibraryError("VAEDecoderF32 needs CUDA tensors; there is no CPU path")
        P, D = self.P, self.P + ".decoder"
        pq = P + ".post_quant_conv.weight"
        if pq not in self._w:      # ddpm.py:843 `1/scale_factor * z` folded into the 1x1 post_quant_conv weights
            self._w[pq] = (self.sd[pq].to(self.dev, F32) / self.scale_factor).permute(0, 2, 3, 1).contiguous()
        x = z.to(self.dev, F32).permute(0, 2, 3, 1).contiguous()
        x = self.conv(self.conv(x, P + ".post_quant_conv"), D + ".conv_in")
        x = self._resnet(D + ".mid.block_1", x)
        x = self._attn(D + ".mid.attn_1", x)
        x = self._resnet(D + ".mid.block_2", x)
        for lvl in reversed(range(self._count(D + ".up."))):
            for i in range(self._count(f"{D}.up.{lvl}.block.")):
                x = self._resnet(f"{D}.up.{lvl}.block.{i}", x)
            if (f"{D}.up.{lvl}.upsample.conv.weight") in self.sd:
                x = self.conv(x, f"{D}.up.{lvl}.upsample.conv", up=1)
        x = self.conv(self.gn(x, D + ".norm_out", 1e-6, True), D + ".conv_out")
        return x.permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def decode_u8(self, z):
        """inference.py:85-87 on the fp32 image: uint8 [B,H,W,3]."""
        img = self.decode(z).permute(0, 2, 3, 1).contiguous()       # NHWC fp32, 3 channels
        return ops.image_to_u8(img)
