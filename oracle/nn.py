"""Oracle: SD-2.1 UNet + control adapter step and VAE decoder, plain PyTorch fp32, functional,
driven directly by the reference's flat state_dict.  TEST INFRASTRUCTURE.

Structure is inferred from the state_dict keys (the same layout the reference checkpoint has,
SURVEY.md Appendix A), so no module classes are re-created.  Follows
  model/rdeic.py:174-235 (NoiseEstimator.forward / forward_unconditional),
  model/rdeic.py:487-598 and ldm/modules/diffusionmodules/openaimodel.py:162-274 (ResBlock),
  openaimodel.py:90-152 (Upsample / Downsample), openaimodel.py:73-88 (TimestepEmbedSequential),
  ldm/modules/attention.py:49-72,153-203,255-350 (GEGLU, CrossAttention, BasicTransformerBlock,
  SpatialTransformer), ldm/modules/diffusionmodules/util.py:161-181,224 (timestep_embedding,
  GroupNorm32), ldm/modules/diffusionmodules/model.py:92-205,580-686 (VAE ResnetBlock, AttnBlock,
  Decoder), ldm/models/diffusion/ddpm.py:835-844 + ldm/models/autoencoder.py:97-100
  (decode_first_stage).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def find_denominator(number: int, start: int) -> int:
    """model/rdeic.py:464-471."""
    if start >= number:
        return number
    while start != 0:
        if number % start == 0:
            return start
        start -= 1
    return 1


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """util.py:161-181."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd: SD, p: str, x, eps: float):
    c = x.shape[1]
    return F.group_norm(x.float(), find_denominator(c, 32), sd[p + ".weight"], sd[p + ".bias"], eps)


def _conv(sd: SD, p: str, x, stride: int = 1):
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd.get(p + ".bias"), stride=stride, padding=w.shape[-1] // 2)


def _lin(sd: SD, p: str, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def time_embed(sd: SD, p: str, t_emb):
    """openaimodel.py:539-543 / rdeic.py:323-328: Linear, SiLU, Linear."""
    return _lin(sd, p + ".2", F.silu(_lin(sd, p + ".0", t_emb)))


def resblock(sd: SD, p: str, x, emb):
    """openaimodel.py:249-274 (no up/down, no scale-shift norm)."""
    h = _conv(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)))
    emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb))
    h = h + emb_out[:, :, None, None]
    h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)))
    if (p + ".skip_connection.weight") in sd:
        x = _conv(sd, p + ".skip_connection", x)
    return x + h


def cross_attention(sd: SD, p: str, x, context, heads: int):
    """attention.py:171-203 (fp32 logits, softmax over keys)."""
    q = _lin(sd, p + ".to_q", x)
    ctx = x if context is None else context
    k = _lin(sd, p + ".to_k", ctx)
    v = _lin(sd, p + ".to_v", ctx)
    b, n, inner = q.shape
    d = inner // heads
    sp = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)
    q, k, v = sp(q), sp(k), sp(v)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    sim = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", sim, v).permute(0, 2, 1, 3).reshape(b, n, inner)
    return _lin(sd, p + ".to_out.0", out)


def transformer_block(sd: SD, p: str, x, context, heads: int):
    """attention.py:281-285."""
    ln = lambda name, t: F.layer_norm(t, (t.shape[-1],), sd[f"{p}.{name}.weight"], sd[f"{p}.{name}.bias"], 1e-5)
    x = cross_attention(sd, p + ".attn1", ln("norm1", x), None, heads) + x
    x = cross_attention(sd, p + ".attn2", ln("norm2", x), context, heads) + x
    h = _lin(sd, p + ".ff.net.0.proj", ln("norm3", x))
    a, gate = h.chunk(2, dim=-1)
    x = _lin(sd, p + ".ff.net.2", a * F.gelu(gate)) + x
    return x


def spatial_transformer(sd: SD, p: str, x, context, d_head_cfg: int, is_control: bool):
    """attention.py:331-350 with use_linear=True, depth 1."""
    b, c, h, w = x.shape
    d_head = find_denominator(c, d_head_cfg) if is_control else d_head_cfg   # rdeic.py:372 vs openaimodel.py:592
    heads = c // d_head
    x_in = x
    x = _gn(sd, p + ".norm", x, 1e-6)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    x = _lin(sd, p + ".proj_in", x)
    x = transformer_block(sd, p + ".transformer_blocks.0", x, context, heads)
    x = _lin(sd, p + ".proj_out", x)
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return x + x_in


def _block(sd: SD, p: str, x, emb, context, d_head_cfg: int, is_control: bool):
    """TimestepEmbedSequential (openaimodel.py:79-88): walk children .0, .1, .2 by key shape."""
    i = 0
    while True:
        q = f"{p}.{i}"
        if (q + ".in_layers.0.weight") in sd:
            x = resblock(sd, q, x, emb)
        elif (q + ".norm.weight") in sd and (q + ".proj_in.weight") in sd:
            x = spatial_transformer(sd, q, x, context, d_head_cfg, is_control)
        elif (q + ".op.weight") in sd:
            x = _conv(sd, q + ".op", x, stride=2)                         # Downsample, openaimodel.py:150
        elif (q + ".conv.weight") in sd:
            x = _conv(sd, q + ".conv", F.interpolate(x, scale_factor=2, mode="nearest"))  # Upsample :106-113
        elif (q + ".weight") in sd and sd[q + ".weight"].dim() == 4:
            x = _conv(sd, q, x)                                           # plain conv (input_blocks.0.0)
        else:
            break
        i += 1
    return x


def _count(sd: SD, prefix: str) -> int:
    idx = set()
    for k in sd:
        if k.startswith(prefix):
            idx.add(int(k[len(prefix):].split(".")[0]))
    return max(idx) + 1 if idx else 0


def noise_estimator_forward(sd: SD, x, guide_hint, t, context, model_channels: int = 320, base_d_head: int = 64,
                            ctrl_d_head: int = 16, control_scale: float = 1.0, unconditional: bool = False):
    """model/rdeic.py:174-212 (and :214-235 when `unconditional`).  `sd` is the full RDEIC
    state_dict (keys `model.diffusion_model.*`, `control_model.*`)."""
    B_ = "model.diffusion_model"
    C_ = "control_model.control_model"
    t_emb = timestep_embedding(t, model_channels)
    emb_base = time_embed(sd, B_ + ".time_embed", t_emb)
    n_in = _count(sd, B_ + ".input_blocks.")
    n_out = _count(sd, B_ + ".output_blocks.")
    h_base = x.float()
    hs_base = []
    if unconditional:
        for i in range(n_in):
            h_base = _block(sd, f"{B_}.input_blocks.{i}", h_base, emb_base, context, base_d_head, False)
            hs_base.append(h_base)
        h_base = _block(sd, f"{B_}.middle_block", h_base, emb_base, context, base_d_head, False)
        for i in range(n_out):
            h_base = torch.cat([h_base, hs_base.pop()], dim=1)
            h_base = _block(sd, f"{B_}.output_blocks.{i}", h_base, emb_base, context, base_d_head, False)
    else:
        emb = time_embed(sd, C_ + ".time_embed", t_emb)
        scales = sd["control_model.scale_list"] * control_scale           # rdeic.py:185
        si = 0
        h_ctr = torch.cat((h_base, guide_hint), dim=1)
        hs_ctr = []
        for i in range(n_in):
            h_base = _block(sd, f"{B_}.input_blocks.{i}", h_base, emb_base, context, base_d_head, False)
            h_ctr = _block(sd, f"{C_}.input_blocks.{i}", h_ctr, emb, context, ctrl_d_head, True)
            h_base = h_base + _conv(sd, f"control_model.enc_zero_convs_out.{i}.0", h_ctr) * scales[si]
            si += 1
            hs_base.append(h_base)
            hs_ctr.append(h_ctr)
        h_base = _block(sd, f"{B_}.middle_block", h_base, emb_base, context, base_d_head, False)
        h_ctr = _block(sd, f"{C_}.middle_block", h_ctr, emb, context, ctrl_d_head, True)
        h_base = h_base + _conv(sd, "control_model.middle_block_out.0", h_ctr) * scales[si]
        si += 1
        for i in range(n_out):
            h_base = h_base + _conv(sd, f"control_model.dec_zero_convs_out.{i}.0", hs_ctr.pop()) * scales[si]
            si += 1
            h_base = torch.cat([h_base, hs_base.pop()], dim=1)
            h_base = _block(sd, f"{B_}.output_blocks.{i}", h_base, emb_base, context, base_d_head, False)
    h = F.silu(_gn(sd, B_ + ".out.0", h_base, 1e-5))
    return _conv(sd, B_ + ".out.2", h)


# ---------------------------------------------------------------------------------------------
# VAE decoder
# ---------------------------------------------------------------------------------------------
def _vae_resnet(sd: SD, p: str, x):
    """model.py:128-151 (temb None)."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, 1e-6)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, 1e-6)))
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x)
    return x + h


def _vae_attn(sd: SD, p: str, x):
    """model.py:181-205."""
    h_ = _gn(sd, p + ".norm", x, 1e-6)
    q, k, v = _conv(sd, p + ".q", h_), _conv(sd, p + ".k", h_), _conv(sd, p + ".v", h_)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, p + ".proj_out", h_)


def vae_decode(sd: SD, z, scale_factor: float = 0.18215, prefix: str = "first_stage_model"):
    """ddpm.py:843 (z / scale_factor) -> autoencoder.py:98-99 (post_quant_conv, decoder) ->
    model.py:653-686."""
    D = prefix + ".decoder"
    z = 1.0 / scale_factor * z
    z = _conv(sd, prefix + ".post_quant_conv", z)
    h = _conv(sd, D + ".conv_in", z)
    h = _vae_resnet(sd, D + ".mid.block_1", h)
    h = _vae_attn(sd, D + ".mid.attn_1", h)
    h = _vae_resnet(sd, D + ".mid.block_2", h)
    n_levels = _count(sd, D + ".up.")
    for lvl in reversed(range(n_levels)):
        nb = _count(sd, f"{D}.up.{lvl}.block.")
        for i in range(nb):
            h = _vae_resnet(sd, f"{D}.up.{lvl}.block.{i}", h)
        if (f"{D}.up.{lvl}.upsample.conv.weight") in sd:
            h = _conv(sd, f"{D}.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    h = F.silu(_gn(sd, D + ".norm_out", h, 1e-6))
    return _conv(sd, D + ".conv_out", h)


def vae_encode_hc(sd: SD, x, prefix: str = "first_stage_model"):
    """autoencoder.py:91-95 encode_hc -> model.py:551-577 Encoder.forward_hc: returns `c`, the feature
    map after norm_out + swish (the posterior is unused by model/rdeic.py:660-663)."""
    E = prefix + ".encoder"
    h = _conv(sd, E + ".conv_in", x)
    n_levels = _count(sd, E + ".down.")
    for lvl in range(n_levels):
        for i in range(_count(sd, f"{E}.down.{lvl}.block.")):
            h = _vae_resnet(sd, f"{E}.down.{lvl}.block.{i}", h)
        if (f"{E}.down.{lvl}.downsample.conv.weight") in sd:
            # model.py:82-84: pad (0,1,0,1), stride-2 conv, no padding
            w, b = sd[f"{E}.down.{lvl}.downsample.conv.weight"], sd[f"{E}.down.{lvl}.downsample.conv.bias"]
            h = F.conv2d(F.pad(h, (0, 1, 0, 1), mode="constant", value=0), w, b, stride=2, padding=0)
    h = _vae_resnet(sd, E + ".mid.block_1", h)
    h = _vae_attn(sd, E + ".mid.attn_1", h)
    h = _vae_resnet(sd, E + ".mid.block_2", h)
    return F.silu(_gn(sd, E + ".norm_out", h, 1e-6))


def to_uint8(x: torch.Tensor) -> torch.Tensor:
    """inference.py:85-87: ((x+1)/2).clamp(0,1) -> b h w c * 255 -> clip -> uint8."""
    x = ((x + 1) / 2).clamp(0, 1)
    x = (x.permute(0, 2, 3, 1) * 255).numpy().clip(0, 255).astype("uint8")
    return torch.from_numpy(x)
