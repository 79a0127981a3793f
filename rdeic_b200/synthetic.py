"""Synthetic checkpoints in the reference's state_dict layout.

There is no network in the build/bench environment, so neither the SD-2.1 nor the RDEIC
checkpoint can be fetched; benchmarks and parity tests run on seeded random-init weights of the
exact architecture (BASELINE.json `north_star`).  `state_dict_spec` walks the architecture the way
the reference constructors do (openaimodel.py:538-751, rdeic.py:92-165,323-462,
ldm/modules/diffusionmodules/model.py:605-651) and lists every decode-path tensor with its shape;
tests/golden/state_dict_keys.json pins that list against the reference's own `state_dict()`.

Zero-initialised modules (`zero_module`, openaimodel.py:228,750; attention.py:328; rdeic.py:171)
would make `apply_model` return exactly 0, so they get N(0, 0.02) like every other tensor gets a
fan-in scaled uniform (PyTorch's default conv/linear init).
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Mapping, Tuple

import torch

from .engine import find_denominator

Spec = Tuple[str, Tuple[int, ...], str]   # (key, shape, kind): kind in w | zw | b | g | beta | buf


def _conv(p: str, cout: int, cin: int, k: int, zero: bool = False, bias: bool = True) -> List[Spec]:
    out = [(p + ".weight", (cout, cin, k, k), "zw" if zero else "w")]
    if bias:
        out.append((p + ".bias", (cout,), "zb" if zero else "b"))
    return out


def _lin(p: str, cout: int, cin: int, zero: bool = False, bias: bool = True) -> List[Spec]:
    out = [(p + ".weight", (cout, cin), "zw" if zero else "w")]
    if bias:
        out.append((p + ".bias", (cout,), "zb" if zero else "b"))
    return out


def _norm(p: str, c: int) -> List[Spec]:
    return [(p + ".weight", (c,), "g"), (p + ".bias", (c,), "beta")]


def _res(p: str, cin: int, cout: int, emb: int) -> List[Spec]:
    s = _norm(p + ".in_layers.0", cin) + _conv(p + ".in_layers.2", cout, cin, 3) + _lin(p + ".emb_layers.1", cout, emb)
    s += _norm(p + ".out_layers.0", cout) + _conv(p + ".out_layers.3", cout, cout, 3, zero=True)
    if cin != cout:
        s += _conv(p + ".skip_connection", cout, cin, 1)
    return s


def _st(p: str, c: int, ctx: int) -> List[Spec]:
    t = p + ".transformer_blocks.0"
    s = _norm(p + ".norm", c) + _lin(p + ".proj_in", c, c)
    for n in ("to_q", "to_k", "to_v"):
        s += _lin(f"{t}.attn1.{n}", c, c, bias=False)
    s += _lin(t + ".attn1.to_out.0", c, c)
    s += _lin(t + ".ff.net.0.proj", 8 * c, c) + _lin(t + ".ff.net.2", c, 4 * c)
    s += _lin(t + ".attn2.to_q", c, c, bias=False) + _lin(t + ".attn2.to_k", c, ctx, bias=False)
    s += _lin(t + ".attn2.to_v", c, ctx, bias=False) + _lin(t + ".attn2.to_out.0", c, c)
    s += _norm(t + ".norm1", c) + _norm(t + ".norm2", c) + _norm(t + ".norm3", c)
    s += _lin(p + ".proj_out", c, c, zero=True)
    return s


def _unet(P: str, cfg: Mapping, *, width: int, emb_dim: int, first_in: int, decoder: bool):
    """Returns (specs, encoder output channels per input block, mid channels, decoder out channels)."""
    mult_list = list(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    nrb = [nrb] * len(mult_list) if isinstance(nrb, int) else list(nrb)
    attn_res = list(cfg["attention_resolutions"])
    ctx = int(cfg["context_dim"])
    mc_full = emb_dim // 4
    s: List[Spec] = _lin(P + ".time_embed.0", emb_dim, mc_full) + _lin(P + ".time_embed.2", emb_dim, emb_dim)
    s += _conv(P + ".input_blocks.0.0", width, first_in, 3)
    enc_out = [width]
    chans = [width]
    ch, ds, idx = width, 1, 1
    for level, mult in enumerate(mult_list):
        for _ in range(nrb[level]):
            s += _res(f"{P}.input_blocks.{idx}.0", ch, mult * width, emb_dim)
            ch = mult * width
            if ds in attn_res:
                s += _st(f"{P}.input_blocks.{idx}.1", ch, ctx)
            enc_out.append(ch)
            chans.append(ch)
            idx += 1
        if level != len(mult_list) - 1:
            s += _conv(f"{P}.input_blocks.{idx}.0.op", ch, ch, 3)
            enc_out.append(ch)
            chans.append(ch)
            idx += 1
            ds *= 2
    s += _res(P + ".middle_block.0", ch, ch, emb_dim) + _st(P + ".middle_block.1", ch, ctx) + _res(P + ".middle_block.2", ch, ch, emb_dim)
    mid = ch
    dec_out: List[int] = []
    if decoder:
        oi = 0
        for level, mult in list(enumerate(mult_list))[::-1]:
            for i in range(nrb[level] + 1):
                ich = chans.pop()
                s += _res(f"{P}.output_blocks.{oi}.0", ch + ich, width * mult, emb_dim)
                ch = width * mult
                j = 1
                if ds in attn_res:
                    s += _st(f"{P}.output_blocks.{oi}.{j}", ch, ctx)
                    j += 1
                if level and i == nrb[level]:
                    s += _conv(f"{P}.output_blocks.{oi}.{j}.conv", ch, ch, 3)
                    ds //= 2
                dec_out.append(ch)
                oi += 1
        s += _norm(P + ".out.0", ch) + _conv(P + ".out.2", int(cfg["out_channels"]), width, 3, zero=True)
    return s, enc_out, mid, dec_out


def _vae_decoder(P: str, dd: Mapping) -> List[Spec]:
    ch, ch_mult, nrb, zc = int(dd["ch"]), list(dd["ch_mult"]), int(dd["num_res_blocks"]), int(dd["z_channels"])
    D = P + ".decoder"
    embed_dim = zc  # rdeic.yaml embed_dim == z_channels
    s: List[Spec] = _conv(P + ".post_quant_conv", zc, embed_dim, 1)
    block_in = ch * ch_mult[-1]
    s += _conv(D + ".conv_in", block_in, zc, 3)

    def rb(p, cin, cout):
        r = _norm(p + ".norm1", cin) + _conv(p + ".conv1", cout, cin, 3) + _norm(p + ".norm2", cout) + _conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            r += _conv(p + ".nin_shortcut", cout, cin, 1)
        return r

    s += rb(D + ".mid.block_1", block_in, block_in)
    s += _norm(D + ".mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        s += _conv(f"{D}.mid.attn_1.{n}", block_in, block_in, 1)
    s += rb(D + ".mid.block_2", block_in, block_in)
    for lvl in reversed(range(len(ch_mult))):
        block_out = ch * ch_mult[lvl]
        for i in range(nrb + 1):
            s += rb(f"{D}.up.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            s += _conv(f"{D}.up.{lvl}.upsample.conv", block_in, block_in, 3)
    s += _norm(D + ".norm_out", block_in) + _conv(D + ".conv_out", int(dd["out_ch"]), block_in, 3)
    return s


def _vae_encoder(P: str, dd: Mapping) -> List[Spec]:
    """ldm/modules/diffusionmodules/model.py:455-522 Encoder + autoencoder.py:37 quant_conv (106 + 2 keys)."""
    ch, ch_mult, nrb, zc = int(dd["ch"]), list(dd["ch_mult"]), int(dd["num_res_blocks"]), int(dd["z_channels"])
    E = P + ".encoder"
    s: List[Spec] = _conv(E + ".conv_in", ch, int(dd["in_channels"]), 3)

    def rb(p, cin, cout):
        r = _norm(p + ".norm1", cin) + _conv(p + ".conv1", cout, cin, 3) + _norm(p + ".norm2", cout) + _conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            r += _conv(p + ".nin_shortcut", cout, cin, 1)
        return r

    block_in = ch
    for lvl, mult in enumerate(ch_mult):
        block_out = ch * mult
        for i in range(nrb):
            s += rb(f"{E}.down.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != len(ch_mult) - 1:
            s += _conv(f"{E}.down.{lvl}.downsample.conv", block_in, block_in, 3)
    s += rb(E + ".mid.block_1", block_in, block_in)
    s += _norm(E + ".mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        s += _conv(f"{E}.mid.attn_1.{n}", block_in, block_in, 1)
    s += rb(E + ".mid.block_2", block_in, block_in)
    zz = 2 * zc if dd.get("double_z", True) else zc
    s += _norm(E + ".norm_out", block_in) + _conv(E + ".conv_out", zz, block_in, 3)
    s += _conv(P + ".quant_conv", 2 * zc, zz, 1)
    return s


def state_dict_spec(params: Mapping, encoder: bool = False) -> List[Spec]:
    """Every decode-path tensor of the RDEIC checkpoint for the given `params` block of
    configs/model/rdeic.yaml: base UNet, control adapter + zero convs, VAE decoder; with `encoder`
    also the sender side's VAE encoder (appended last, so the decode-path tensors keep their values)."""
    up = dict(params["unet_config"]["params"])
    cp = dict(params["control_stage_config"]["params"])
    dd = dict(params["first_stage_config"]["params"]["ddconfig"])
    mc = int(up["model_channels"])
    base, b_enc, b_mid, b_dec = _unet("model.diffusion_model", up, width=mc, emb_dim=4 * mc,
                                      first_in=int(up["in_channels"]), decoder=True)
    cmc = int(cp["model_channels"])
    cw = int(cmc * float(cp.get("control_model_ratio", 1.0)))
    ctrl, c_enc, c_mid, _ = _unet("control_model.control_model", cp, width=cw, emb_dim=4 * cmc,
                                  first_in=int(cp["in_channels"]) + int(cp["hint_channels"]), decoder=False)
    z: List[Spec] = []
    for i, (co, bo) in enumerate(zip(c_enc, b_enc)):                       # rdeic.py:155-158
        z += _conv(f"control_model.enc_zero_convs_out.{i}.0", bo, co, 1, zero=True)
    z += _conv("control_model.middle_block_out.0", b_mid, c_mid, 1, zero=True)   # rdeic.py:145
    z += _conv("control_model.dec_zero_convs_out.0.0", b_mid, c_enc[-1], 1, zero=True)   # rdeic.py:147-149
    for i in range(1, len(c_enc)):                                          # rdeic.py:150-153
        z += _conv(f"control_model.dec_zero_convs_out.{i}.0", b_dec[i - 1], c_enc[-(i + 1)], 1, zero=True)
    z.append(("control_model.scale_list", (2 * len(c_enc) + 1,), "buf"))
    out = base + ctrl + z + _vae_decoder("first_stage_model", dd)
    return out + _vae_encoder("first_stage_model", dd) if encoder else out


# ---------------------------------------------------------------------------------------------
# preprocess_model: the learned compressor (model/compression.py:10-50)
# ---------------------------------------------------------------------------------------------
def _cres(p: str, cin: int, cout: int) -> List[Spec]:
    """model/layers/res_blk.py:65-96 ResidualBlock."""
    s = _conv(p + ".conv1", cout, cin, 3) + _conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        s += _conv(p + ".adaptor", cout, cin, 1)
    return s


def _cres_up(p: str, cin: int, cout: int) -> List[Spec]:
    """res_blk.py:39-63 ResidualBlockUpsample (sub-pixel 1x1 convs, model/layers/conv.py:7-10)."""
    return _conv(p + ".subpel_conv.0", 4 * cout, cin, 1) + _conv(p + ".conv", cout, cout, 3) + \
        _conv(p + ".upsample.0", 4 * cout, cin, 1)


def _cres_down(p: str, cin: int, cout: int) -> List[Spec]:
    """res_blk.py:6-37 ResidualBlockWithStride."""
    return _conv(p + ".conv1", cout, cin, 3) + _conv(p + ".conv2", cout, cout, 3) + _conv(p + ".downsample", cout, cin, 1)


def compression_state_dict_spec(pp: Mapping, prefix: str = "preprocess_model.") -> List[Spec]:
    """Every tensor of `Compression(in_nc, out_nc, N, M, slice_num, slice_ch, codebook_size)` in the
    order of the reference constructor (model/compression.py:11-50, compression_modules.py:7-104)."""
    in_nc, out_nc, N, M = int(pp["in_nc"]), int(pp["out_nc"]), int(pp["N"]), int(pp["M"])
    sc = [int(c) for c in pp["slice_ch"]]
    P = prefix
    s: List[Spec] = []
    ga = P + "encoder.g_a."
    s += _cres(ga + "0", in_nc, M)
    for i in (1, 2, 3):
        s += _cres(ga + str(i), M, M)
    s += _cres_down(ga + "4", M, M)
    for i in (5, 6, 7):
        s += _cres(ga + str(i), M, M)
    s += _conv(ga + "8", M, M, 3)
    he = P + "hyper_enc.hyper_enc."
    s += _cres(he + "0", M, N) + _cres(he + "1", N, N) + _cres_down(he + "2", N, N) + _cres_down(he + "3", N, N)
    hd = P + "hyper_dec.hyper_dec."
    s += _cres_up(hd + "0", N, M) + _cres_up(hd + "1", M, M) + _cres(hd + "2", M, M * 3 // 2) + _cres(hd + "3", M * 3 // 2, 2 * M)
    gs = P + "decoder.g_s."
    s += _conv(gs + "0", M, M, 3)
    for i in (1, 2, 3):
        s += _cres(gs + str(i), M, M)
    s += _cres_up(gs + "4", M, M)
    for i in (5, 6, 7, 8):
        s += _cres(gs + str(i), M, M)
    s += _conv(P + "out", out_nc, M, 3)
    for i, c in enumerate(sc):
        s += _conv(f"{P}local_context.{i}", 2 * c, c, 5)
    for i, c in enumerate(sc):
        if i:
            q = f"{P}channel_context.{i}.fushion."
            s += _conv(q + "0", 224, sum(sc[:i]), 5) + _conv(q + "2", 128, 224, 5) + _conv(q + "4", 2 * c, 128, 5)

    def ep(q: str, cin: int, cout: int) -> List[Spec]:
        return _conv(q + "0", cout * 5 // 3, cin, 1) + _conv(q + "2", cout * 4 // 3, cout * 5 // 3, 1) + \
            _conv(q + "4", cout, cout * 4 // 3, 1)

    for i, c in enumerate(sc):
        s += ep(f"{P}entropy_parameters_anchor.{i}.fusion.", 2 * M + (2 * c if i else 0), 2 * c)
    for i, c in enumerate(sc):
        s += ep(f"{P}entropy_parameters_nonanchor.{i}.fusion.", 2 * M + (4 * c if i else 2 * c), 2 * c)
    K = int(pp["codebook_size"])
    s.append((P + "quantize.embedding.weight", (K, N), "code"))
    s.append((P + "quantize.embed_prob", (K,), "prob"))
    return s


def make_compression_state_dict(pp: Mapping, seed: int = 232, device="cpu", prefix: str = "preprocess_model.") -> Dict[str, torch.Tensor]:
    """Seeded compressor checkpoint.  Weights are variance-preserving (uniform +-sqrt(3/fan_in)) and
    the VQ codebook is N(0,1) so that, unlike PyTorch's shrinking default init, the entropy
    parameters spread over the scale table and the symbols are not all zero."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in compression_state_dict_spec(pp, prefix):
        if kind == "w":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * math.sqrt(3.0 / fan_in)
        elif kind == "b":
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * 0.1
        elif kind == "code":
            t = torch.randn(shape, generator=g, device=dev)
        elif kind == "prob":
            t = torch.zeros(shape, device=dev)
        else:
            raise ValueError(kind)
        sd[key] = t.float()
    return sd


def make_state_dict(params: Mapping, seed: int = 231, device="cpu", control_scale: float = 1.0,
                    encoder: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded random checkpoint (fp32) with the reference layout."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_spec(params, encoder=encoder):
        if kind in ("w", "b"):
            fan_in = 1
            if kind == "w":
                for d in shape[1:]:
                    fan_in *= d
                bound = 1.0 / math.sqrt(fan_in)
            else:
                wshape = sd[key[:-len(".bias")] + ".weight"].shape
                for d in wshape[1:]:
                    fan_in *= d
                bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * bound
        elif kind in ("zw", "zb"):
            t = torch.randn(shape, generator=g, device=dev) * 0.02
        elif kind == "g":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=dev)
        elif kind == "beta":
            t = 0.1 * torch.randn(shape, generator=g, device=dev)
        elif kind == "buf":
            t = torch.ones(shape, device=dev) * control_scale
        else:
            raise ValueError(kind)
        sd[key] = t.float()
    return sd
