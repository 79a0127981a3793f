"""Thin torch-tensor wrappers over the C ABI (include/rdeic_b200.h).

Every function takes CUDA tensors, passes raw pointers + the current torch stream to
librdeic_b200.so and returns torch tensors.  No arithmetic happens in PyTorch here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ConvParams
from ._lib import check as _check

BF16 = torch.bfloat16

# Kernel-launch accounting (bench.py's `gpu_launches`) and optional per-GEMM timing hooks.
LAUNCHES = 0          # kernels launched by this module (graph replays add their node count)
GEMM_PROFILE = None   # when a list: conv_gemm appends (start_event, end_event, flops)


_KERNELS_PER_CALL = {"rdeic_groupnorm_nhwc": 2, "rdeic_groupnorm_from_stats": 2, "rdeic_vq_quant": 3,
                     "rdeic_gn_silu_conv3x3_tail": 2, "rdeic_groupnorm_nhwc(small)": 1}


def check(status: int, what: str) -> None:
    global LAUNCHES
    _check(status, what)
    LAUNCHES += _KERNELS_PER_CALL.get(what, 1)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.RdeicLibraryError(
            f"{name}: expected a CUDA tensor (this path has no CPU fallback), got {t.device}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# ---------------------------------------------------------------------------------------------
# entropy front end
# ---------------------------------------------------------------------------------------------
def ckbd_mask(y: torch.Tensor, which: int) -> torch.Tensor:
    y = _need(y, torch.float32, "ckbd_mask")
    B, Cc, H, W = y.shape
    out = torch.empty_like(y)
    if y.numel() == 0:
        return out
    check(_lib.load().rdeic_ckbd_mask(_ptr(y), _ptr(out), B, Cc, H, W, which, _stream()), "rdeic_ckbd_mask")
    return out


def ckbd_split(y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    y = _need(y, torch.float32, "ckbd_split")
    B, Cc, H, W = y.shape
    a, n = torch.empty_like(y), torch.empty_like(y)
    if y.numel() == 0:
        return a, n
    check(_lib.load().rdeic_ckbd_split(_ptr(y), _ptr(a), _ptr(n), B, Cc, H, W, _stream()), "rdeic_ckbd_split")
    return a, n


def ckbd_merge(a: torch.Tensor, n: torch.Tensor) -> torch.Tensor:
    a = _need(a, torch.float32, "ckbd_merge")
    n = _need(n, torch.float32, "ckbd_merge")
    if a.shape != n.shape:
        raise RuntimeError(f"ckbd_merge: shape mismatch {tuple(a.shape)} vs {tuple(n.shape)}")
    out = torch.empty_like(a)
    check(_lib.load().rdeic_ckbd_merge(_ptr(a), _ptr(n), _ptr(out), a.numel(), _stream()), "rdeic_ckbd_merge")
    return out


def ckbd_squeeze(y: torch.Tensor, which: int) -> torch.Tensor:
    y = _need(y, torch.float32, "ckbd_squeeze")
    B, Cc, H, W = y.shape
    out = torch.empty((B, Cc, H, W // 2), dtype=torch.float32, device=y.device)
    if W % 2 == 0 and y.numel() == 0:
        return out
    check(_lib.load().rdeic_ckbd_squeeze(_ptr(y), _ptr(out), B, Cc, H, W, which, _stream()), "rdeic_ckbd_squeeze")
    return out


def ckbd_unsqueeze(s: torch.Tensor, which: int) -> torch.Tensor:
    s = _need(s, torch.float32, "ckbd_unsqueeze")
    B, Cc, H, Wh = s.shape
    out = torch.empty((B, Cc, H, Wh * 2), dtype=torch.float32, device=s.device)
    if s.numel() == 0:
        return out
    check(_lib.load().rdeic_ckbd_unsqueeze(_ptr(s), _ptr(out), B, Cc, H, Wh, which, _stream()), "rdeic_ckbd_unsqueeze")
    return out


def quantize_symbols(x: torch.Tensor, means: Optional[torch.Tensor]) -> torch.Tensor:
    x = _need(x, torch.float32, "quantize_symbols")
    if means is not None:
        means = _need(means, torch.float32, "quantize_symbols").expand_as(x).contiguous()
    out = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    check(_lib.load().rdeic_quantize_symbols(_ptr(x), _ptr(means), _ptr(out), x.numel(), _stream()),
          "rdeic_quantize_symbols")
    return out


def dequantize(sym: torch.Tensor, means: torch.Tensor) -> torch.Tensor:
    sym = _need(sym, torch.int32, "dequantize")
    means = _need(means, torch.float32, "dequantize")
    out = torch.empty(sym.shape, dtype=torch.float32, device=sym.device)
    check(_lib.load().rdeic_dequantize(_ptr(sym), _ptr(means), _ptr(out), sym.numel(), _stream()), "rdeic_dequantize")
    return out


def build_indexes(scales: torch.Tensor, table: torch.Tensor, lower_bound: float) -> torch.Tensor:
    scales = _need(scales, torch.float32, "build_indexes")
    table = _need(table, torch.float32, "build_indexes")
    out = torch.empty(scales.shape, dtype=torch.int32, device=scales.device)
    check(_lib.load().rdeic_build_indexes(_ptr(scales), _ptr(table), table.numel(), lower_bound, _ptr(out),
                                          scales.numel(), _stream()), "rdeic_build_indexes")
    return out


def ckbd_squeeze_indexes(scales, means, table, lower_bound: float, which: int):
    scales = _need(scales, torch.float32, "ckbd_squeeze_indexes")
    means = _need(means, torch.float32, "ckbd_squeeze_indexes")
    table = _need(table, torch.float32, "ckbd_squeeze_indexes")
    B, Cc, H, W = scales.shape
    means_sq = torch.empty((B, Cc, H, W // 2), dtype=torch.float32, device=scales.device)
    idx = torch.empty((B, Cc, H, W // 2), dtype=torch.int32, device=scales.device)
    check(_lib.load().rdeic_ckbd_squeeze_indexes(_ptr(scales), _ptr(means), _ptr(table), table.numel(), lower_bound,
                                                 _ptr(means_sq), _ptr(idx), B, Cc, H, W, which, _stream()),
          "rdeic_ckbd_squeeze_indexes")
    return means_sq, idx


def ckbd_encode_phase(y, scales, means, table, lower_bound: float, which: int, sym_out=None, idx_out=None):
    """`sym_out` / `idx_out`: optional flat int32 windows of a stream-order staging buffer."""
    y = _need(y, torch.float32, "ckbd_encode_phase")
    scales = _need(scales, torch.float32, "ckbd_encode_phase")
    means = _need(means, torch.float32, "ckbd_encode_phase")
    table = _need(table, torch.float32, "ckbd_encode_phase")
    B, Cc, H, W = y.shape
    n = B * Cc * H * (W // 2)
    for t in (sym_out, idx_out):
        if t is not None and (t.dtype != torch.int32 or t.numel() != n or not t.is_contiguous()):
            raise ValueError("ckbd_encode_phase: staging windows must be contiguous int32 of B*C*H*W/2 elements")
    sym = sym_out.view(B, Cc, H, W // 2) if sym_out is not None else torch.empty((B, Cc, H, W // 2), dtype=torch.int32, device=y.device)
    idx = idx_out.view(B, Cc, H, W // 2) if idx_out is not None else torch.empty_like(sym)
    y_hat = torch.empty_like(y)
    check(_lib.load().rdeic_ckbd_encode_phase(_ptr(y), _ptr(scales), _ptr(means), _ptr(table), table.numel(),
                                              lower_bound, _ptr(sym), _ptr(idx), _ptr(y_hat), B, Cc, H, W, which,
                                              _stream()), "rdeic_ckbd_encode_phase")
    return sym, idx, y_hat


def ckbd_decode_phase(sym: torch.Tensor, means_sq: torch.Tensor, which: int) -> torch.Tensor:
    sym = _need(sym, torch.int32, "ckbd_decode_phase")
    means_sq = _need(means_sq, torch.float32, "ckbd_decode_phase")
    B, Cc, H, Wh = sym.shape
    out = torch.empty((B, Cc, H, 2 * Wh), dtype=torch.float32, device=sym.device)
    check(_lib.load().rdeic_ckbd_decode_phase(_ptr(sym), _ptr(means_sq), _ptr(out), B, Cc, H, Wh, which, _stream()),
          "rdeic_ckbd_decode_phase")
    return out


def vq_quant(z: torch.Tensor, codebook: torch.Tensor):
    z = _need(z, torch.float32, "vq_quant")
    codebook = _need(codebook, torch.float32, "vq_quant")
    B, D, H, W = z.shape
    K = codebook.shape[0]
    idx = torch.empty((B, H, W), dtype=torch.int64, device=z.device)
    zq = torch.empty_like(z)
    check(_lib.load().rdeic_vq_quant(_ptr(z), _ptr(codebook), _ptr(idx), _ptr(zq), B, D, H * W, K, _stream()),
          "rdeic_vq_quant")
    return zq, idx


def vq_lookup(idx: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    idx = _need(idx, torch.int64, "vq_lookup")
    codebook = _need(codebook, torch.float32, "vq_lookup")
    B, H, W = idx.shape
    K, D = codebook.shape
    out = torch.empty((B, D, H, W), dtype=torch.float32, device=idx.device)
    check(_lib.load().rdeic_vq_lookup(_ptr(idx), _ptr(codebook), _ptr(out), B, D, H * W, K, _stream()), "rdeic_vq_lookup")
    return out


# ---------------------------------------------------------------------------------------------
# sampler updates
# ---------------------------------------------------------------------------------------------
def q_sample(x0, noise, a: float, b: float, out=None):
    x0 = _need(x0, torch.float32, "q_sample")
    noise = _need(noise, torch.float32, "q_sample")
    out = torch.empty_like(x0) if out is None else out
    check(_lib.load().rdeic_q_sample(_ptr(x0), _ptr(noise), _ptr(out), x0.numel(), a, b, _stream()), "rdeic_q_sample")
    return out


def relay_update(x, eps, noise, r, rm1, c1, c2, sigma, eps_uncond=None, guidance_scale=1.0, out=None):
    x = _need(x, torch.float32, "relay_update")
    eps = _need(eps, torch.float32, "relay_update")
    noise = _need(noise, torch.float32, "relay_update")
    if eps_uncond is not None:
        eps_uncond = _need(eps_uncond, torch.float32, "relay_update")
    out = torch.empty_like(x) if out is None else out
    check(_lib.load().rdeic_relay_update(_ptr(x), _ptr(eps), _ptr(eps_uncond), guidance_scale, _ptr(noise), _ptr(out),
                                         x.numel(), r, rm1, c1, c2, sigma, _stream()), "rdeic_relay_update")
    return out


def ddim_update(x, eps, noise, sqrt_one_minus_at, sqrt_at, sqrt_aprev, dir_coef, sigma, eps_uncond=None,
                guidance_scale=1.0):
    x = _need(x, torch.float32, "ddim_update")
    eps = _need(eps, torch.float32, "ddim_update")
    noise = _need(noise, torch.float32, "ddim_update")
    if eps_uncond is not None:
        eps_uncond = _need(eps_uncond, torch.float32, "ddim_update")
    out = torch.empty_like(x)
    pred = torch.empty_like(x)
    check(_lib.load().rdeic_ddim_update(_ptr(x), _ptr(eps), _ptr(eps_uncond), guidance_scale, _ptr(noise), _ptr(out),
                                        _ptr(pred), x.numel(), sqrt_one_minus_at, sqrt_at, sqrt_aprev, dir_coef, sigma,
                                        _stream()), "rdeic_ddim_update")
    return out, pred


# ---------------------------------------------------------------------------------------------
# layout / glue
# ---------------------------------------------------------------------------------------------
def nchw_to_nhwc_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None, ldc: Optional[int] = None, c_off: int = 0):
    src = _need(src, torch.float32, "nchw_to_nhwc_bf16")
    B, Cc, H, W = src.shape
    if dst is None:
        ldc = ldc or Cc
        dst = torch.zeros((B, H, W, ldc), dtype=BF16, device=src.device) if ldc != Cc else \
            torch.empty((B, H, W, ldc), dtype=BF16, device=src.device)
    ldc = dst.shape[-1]
    check(_lib.load().rdeic_nchw_to_nhwc_bf16(_ptr(src), _ptr(dst), B, Cc, H, W, ldc, c_off, _stream()),
          "rdeic_nchw_to_nhwc_bf16")
    return dst


def nhwc_to_nchw_f32(src: torch.Tensor, Cc: Optional[int] = None) -> torch.Tensor:
    """src may be a channel slice of a dense NHWC tensor (its pixel stride is the ldc passed down)."""
    B, H, W, cs = src.shape
    Cc = Cc or cs
    ldc = src.stride(-2)
    if not _is_nhwc_slice(src):
        raise TypeError("nhwc_to_nchw_f32: expected a dense NHWC tensor or a channel slice of one")
    is_f32 = 1 if src.dtype == torch.float32 else 0
    if not is_f32 and src.dtype != BF16:
        raise TypeError("nhwc_to_nchw_f32: expected bf16 or fp32")
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=src.device)
    check(_lib.load().rdeic_nhwc_to_nchw_f32(_ptr(src), is_f32, _ptr(out), B, Cc, H, W, ldc, _stream()),
          "rdeic_nhwc_to_nchw_f32")
    return out


def f32_to_bf16(src: torch.Tensor) -> torch.Tensor:
    src = _need(src, torch.float32, "f32_to_bf16")
    out = torch.empty(src.shape, dtype=BF16, device=src.device)
    check(_lib.load().rdeic_f32_to_bf16(_ptr(src), _ptr(out), src.numel(), _stream()), "rdeic_f32_to_bf16")
    return out


def split_hilo(src: torch.Tensor, dst: Optional[torch.Tensor] = None, off_hi: int = 0, off_lo: Optional[int] = None):
    """src fp32 [..., C] (dense rows) -> bf16 [..., ld] holding hi = bf16(src) in columns [off_hi, off_hi + C) and
    lo = bf16(src - hi) in [off_lo, off_lo + C) (default: right behind hi)."""
    src = _need(src, torch.float32, "split_hilo")
    Cc = src.shape[-1]
    rows = src.numel() // Cc
    off_lo = off_hi + Cc if off_lo is None else off_lo
    if dst is None:
        dst = torch.zeros((*src.shape[:-1], max(off_hi, off_lo) + Cc), dtype=BF16, device=src.device)
    if dst.dtype != BF16 or not dst.is_contiguous() or dst.numel() // dst.shape[-1] != rows:
        raise TypeError("split_hilo: dst must be a contiguous bf16 tensor with the rows of src")
    check(_lib.load().rdeic_split_bf16_hilo(_ptr(src), rows, Cc, Cc, _ptr(dst), dst.shape[-1], off_hi, off_lo, _stream()),
          "rdeic_split_bf16_hilo")
    return dst


def timestep_embedding_f32(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    t = _need(t, torch.int64, "timestep_embedding_f32")
    out = torch.empty((t.shape[0], dim), dtype=torch.float32, device=t.device)
    check(_lib.load().rdeic_timestep_embedding_f32(_ptr(t), _ptr(out), t.shape[0], dim, max_period, _stream()),
          "rdeic_timestep_embedding_f32")
    return out


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    t = _need(t, torch.int64, "timestep_embedding")
    out = torch.empty((t.shape[0], dim), dtype=BF16, device=t.device)
    check(_lib.load().rdeic_timestep_embedding(_ptr(t), _ptr(out), t.shape[0], dim, max_period, _stream()),
          "rdeic_timestep_embedding")
    return out


def silu_bf16(x: torch.Tensor) -> torch.Tensor:
    is_f32 = 1 if x.dtype == torch.float32 else 0
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(_lib.load().rdeic_silu_bf16(_ptr(x), is_f32, _ptr(out), x.numel(), _stream()), "rdeic_silu_bf16")
    return out


def geglu(x: torch.Tensor) -> torch.Tensor:
    rows = x.numel() // x.shape[-1]
    F = x.shape[-1] // 2
    out = torch.empty((*x.shape[:-1], F), dtype=BF16, device=x.device)
    check(_lib.load().rdeic_geglu(_ptr(x), _ptr(out), rows, F, _stream()), "rdeic_geglu")
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    B, H, W, Cc = x.shape
    out = torch.empty((B, 2 * H, 2 * W, Cc), dtype=BF16, device=x.device)
    check(_lib.load().rdeic_upsample2x_nhwc(_ptr(x), _ptr(out), B, H, W, Cc, _stream()), "rdeic_upsample2x_nhwc")
    return out


def pixel_shuffle2(x: torch.Tensor) -> torch.Tensor:
    """nn.PixelShuffle(2) on NHWC bf16 [B,H,W,4C] whose channels are ordered (i, j, c) -> [B,2H,2W,C]."""
    B, H, W, C4 = x.shape
    x = _need(x, BF16, "pixel_shuffle2")
    out = torch.empty((B, 2 * H, 2 * W, C4 // 4), dtype=BF16, device=x.device)
    check(_lib.load().rdeic_pixel_shuffle2_nhwc(_ptr(x), _ptr(out), B, H, W, C4 // 4, _stream()),
          "rdeic_pixel_shuffle2_nhwc")
    return out


def im2col_3x3_s2(x: torch.Tensor, pad_lo: int = 1) -> torch.Tensor:
    """pad_lo = 1: padding 1 on every side; 0: bottom/right only (VAE encoder Downsample)."""
    B, H, W, Cc = x.shape
    Cp = (Cc + 63) // 64 * 64
    out = torch.empty((B * (H // 2) * (W // 2), 9 * Cp), dtype=BF16, device=x.device)
    check(_lib.load().rdeic_im2col_3x3_s2(_ptr(x), _ptr(out), B, H, W, Cc, pad_lo, _stream()), "rdeic_im2col_3x3_s2")
    return out


def softmax_rows(x: torch.Tensor, scale: float) -> torch.Tensor:
    n = x.shape[-1]
    rows = x.numel() // n
    is_f32 = 1 if x.dtype == torch.float32 else 0
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(_lib.load().rdeic_softmax_rows(_ptr(x), is_f32, _ptr(out), rows, n, scale, _stream()), "rdeic_softmax_rows")
    return out


def transpose_bf16(x: torch.Tensor) -> torch.Tensor:
    batch, R, Cc = x.shape
    out = torch.empty((batch, Cc, R), dtype=BF16, device=x.device)
    check(_lib.load().rdeic_transpose_bf16(_ptr(x), _ptr(out), batch, R, Cc, _stream()), "rdeic_transpose_bf16")
    return out


def image_to_u8(x: torch.Tensor) -> torch.Tensor:
    """x: NHWC fp32 [B,H,W,ldc] in [-1,1] -> uint8 [B,H,W,3] (inference.py:85-87)."""
    x = _need(x, torch.float32, "image_to_u8")
    B, H, W, ldc = x.shape
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=x.device)
    check(_lib.load().rdeic_image_to_u8(_ptr(x), _ptr(out), B * H * W, ldc, _stream()), "rdeic_image_to_u8")
    return out


def blend_tiles_u8(tiles: torch.Tensor, origins: torch.Tensor, overlap: int, H: int, W: int) -> torch.Tensor:
    """tiles uint8 [T,th,tw,3], origins int32 [T,2] (y0, x0 in pixels) -> uint8 [H,W,3]."""
    tiles = _need(tiles, torch.uint8, "blend_tiles_u8")
    origins = _need(origins, torch.int32, "blend_tiles_u8")
    T, th, tw, c = tiles.shape
    if c != 3 or tuple(origins.shape) != (T, 2):
        raise ValueError("blend_tiles_u8: expected tiles [T,th,tw,3] and origins [T,2]")
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=tiles.device)
    check(_lib.load().rdeic_blend_tiles_u8(_ptr(tiles), _ptr(origins), T, th, tw, overlap, _ptr(out), H, W, _stream()),
          "rdeic_blend_tiles_u8")
    return out


# ---------------------------------------------------------------------------------------------
# normalisation
# ---------------------------------------------------------------------------------------------
_gn_ws = {}
WS_SLOT = 0     # workspaces are per (device, slot): kernels running concurrently on two streams
                # (base UNet / control adapter overlap) must not share scratch memory


def _gn_workspace(B: int, device) -> torch.Tensor:
    key = (B, str(device), WS_SLOT)
    ws = _gn_ws.get(key)
    if ws is None:
        nbytes = _lib.load().rdeic_groupnorm_workspace_bytes(B, 1, 8)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _gn_ws[key] = ws
    return ws


def groupnorm(x1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool,
              x2: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
              stats1: Optional[torch.Tensor] = None, stats2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x1 [B,H,W,C1] (+ x2 [B,H,W,C2]) bf16 NHWC -> [B,H,W,C1+C2] bf16.  With `stats1` (and `stats2`
    when x2 is given) — the slab statistics the producing conv_gemm emitted — the statistics pass over
    the tensors is skipped."""
    B, H, W, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[-1]
    if x2 is not None and x2.dtype != x1.dtype:
        raise TypeError("groupnorm: both sources must share a dtype")
    out = torch.empty((B, H, W, C1 + C2), dtype=BF16, device=x1.device)
    ws = workspace if workspace is not None else _gn_workspace(B, x1.device)
    if _lib.load().rdeic_groupnorm_is_small(B, H * W, C1, C2, groups):
        # one kernel per GroupNorm for the small tensors of UNet levels 2-3 / the control adapter (fused statistics,
        # if any, are ignored: fold + apply would be two launch latencies)
        check(_lib.load().rdeic_groupnorm_nhwc(_ptr(x1), C1, _ptr(x2), C2, int(x1.dtype == torch.float32), _ptr(gamma),
                                               _ptr(beta), _ptr(out), B, H * W, groups, eps, 1 if silu else 0, _ptr(ws),
                                               _stream()), "rdeic_groupnorm_nhwc(small)")
        return out
    if stats1 is not None and (x2 is None or stats2 is not None):
        check(_lib.load().rdeic_groupnorm_from_stats(_ptr(x1), C1, _ptr(stats1), _ptr(x2), C2, _ptr(stats2),
                                                     int(x1.dtype == torch.float32), _ptr(gamma), _ptr(beta), _ptr(out),
                                                     B, H * W, groups, eps, 1 if silu else 0, _ptr(ws), _stream()),
              "rdeic_groupnorm_from_stats")
        return out
    check(_lib.load().rdeic_groupnorm_nhwc(_ptr(x1), C1, _ptr(x2), C2, int(x1.dtype == torch.float32), _ptr(gamma),
                                           _ptr(beta), _ptr(out), B, H * W, groups, eps, 1 if silu else 0, _ptr(ws),
                                           _stream()), "rdeic_groupnorm_nhwc")
    return out


def groupnorm_f32(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool) -> torch.Tensor:
    """fp32 NHWC in -> fp32 NHWC out (exact exp / division in SiLU): the GroupNorm of the fp32 kernel mode, also the
    producer of the [hi | lo | hi] operands of the split-bf16 convs."""
    x = _need(x, torch.float32, "groupnorm_f32")
    B, H, W, C = x.shape
    out = torch.empty_like(x)
    ws = _gn_workspace(B, x.device)
    check(_lib.load().rdeic_groupnorm_nhwc_f32(_ptr(x), C, None, 0, _ptr(gamma), _ptr(beta), _ptr(out), B, H * W, groups, eps,
                                               1 if silu else 0, _ptr(ws), _stream()), "rdeic_groupnorm_nhwc_f32")
    return out


def split3(src: torch.Tensor) -> torch.Tensor:
    """fp32 [..., C] -> bf16 [..., ceil8(3C)] = [hi | lo | hi] (zero padded): the A operand of a three-pass split-bf16
    product against weights packed as [w_hi | w_hi | w_lo] (`engine.Conv.load(split3=True)`):
    a.w = a_hi.w_hi + a_lo.w_hi + a_hi.w_lo + O(2^-16), fp32 accumulation on the tensor cores."""
    src = _need(src, torch.float32, "split3")
    Cc = src.shape[-1]
    dst = torch.zeros((*src.shape[:-1], (3 * Cc + 7) // 8 * 8), dtype=BF16, device=src.device)
    split_hilo(src, dst, 0, Cc)
    split_hilo(src, dst, 2 * Cc, Cc)
    return dst


def pack_tail_weight(w: torch.Tensor) -> torch.Tensor:
    """conv_out weight OIHW fp32 [n_out <= 4, C, 3, 3] -> bf16 [9 taps, 8 (zero-padded n_out), C] for `gn_silu_conv3x3_tail`."""
    n_out, cin, kh, kw = w.shape
    if (kh, kw) != (3, 3) or n_out > 8:
        raise ValueError("pack_tail_weight: expected a 3x3 kernel with at most 8 output channels")
    out = torch.zeros((9, 8, cin), dtype=BF16, device=w.device)
    out[:, :n_out] = w.float().permute(2, 3, 0, 1).reshape(9, n_out, cin).to(BF16)
    return out.contiguous()


def gn_silu_conv3x3_tail(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int,
                         eps: float, w_tail: torch.Tensor, bias: torch.Tensor, n_out: int, as_uint8: bool,
                         ldo: int = 4) -> torch.Tensor:
    """VAE decoder tail (model.py:683-686 + inference.py:85-87) in one kernel: GroupNorm (statistics from the
    producing conv's epilogue) + SiLU + conv3x3 to n_out <= 4 channels, written as fp32 NHWC [B,H,W,ldo] or,
    with `as_uint8`, as the caller's uint8 HWC image."""
    B, H, W, Cc = x.shape
    _need(x, BF16, "gn_silu_conv3x3_tail")
    ws = _gn_workspace(B, x.device)
    if as_uint8:
        out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=x.device)
        of, ou = None, out
    else:
        out = torch.empty((B, H, W, ldo), dtype=torch.float32, device=x.device)
        of, ou = out, None
    check(_lib.load().rdeic_gn_silu_conv3x3_tail(_ptr(x), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(w_tail), _ptr(bias),
                                                 n_out, _ptr(of), ldo, _ptr(ou), B, H, W, Cc, groups, eps, _ptr(ws),
                                                 _stream()), "rdeic_gn_silu_conv3x3_tail")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(_lib.load().rdeic_layernorm(_ptr(x), int(x.dtype == torch.float32), _ptr(gamma), _ptr(beta), _ptr(out), rows,
                                      Cc, eps, _stream()), "rdeic_layernorm")
    return out


# ---------------------------------------------------------------------------------------------
# tensor-core contractions
# ---------------------------------------------------------------------------------------------
_splitk_ws = {}


def _splitk_workspace(device) -> torch.Tensor:
    """Caller-owned split-K scratch (the library never allocates): 64 MB per device."""
    key = (str(device), WS_SLOT)
    ws = _splitk_ws.get(key)
    if ws is None:
        ws = torch.empty(64 << 20, dtype=torch.uint8, device=device)
        _splitk_ws[key] = ws
    return ws


def pack_conv_weight(w: torch.Tensor, c1: Optional[int] = None) -> torch.Tensor:
    """OIHW (or [out,in]) fp32 -> packed bf16 [n_out, taps*(cp1+cp2)].  c1 splits the input
    channels into two concat sources (c1, cin-c1)."""
    w = _need(w, torch.float32, "pack_conv_weight")
    if w.dim() == 2:
        w = w[:, :, None, None]
    n_out, cin, kh, kw = w.shape
    c1 = cin if c1 is None else c1
    c2 = cin - c1
    cp1, cp2 = (c1 + 63) // 64 * 64, (c2 + 63) // 64 * 64
    out = torch.empty((n_out, kh * kw * (cp1 + cp2)), dtype=BF16, device=w.device)
    check(_lib.load().rdeic_pack_conv_weight(_ptr(w.contiguous()), _ptr(out), n_out, c1, c2, kh, kw, _stream()),
          "rdeic_pack_conv_weight")
    return out


def conv_gemm(a: torch.Tensor, w_packed: torch.Tensor, n_out: int, taps: int, *, a2: Optional[torch.Tensor] = None,
              bias: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None,
              resid: Optional[torch.Tensor] = None, alpha: float = 1.0, act: int = 0, out_f32: bool = False,
              dual: bool = False, out=None, w_batch_stride: int = 0, w_k: int = 0, w_ld: int = 0, tile_n: int = 0,
              split_k: bool = True, act_param: float = 0.0, stats: bool = False, up2: bool = False,
              w2: Optional[torch.Tensor] = None, stride2: bool = False, pad_lo: int = 1):
    """a: NHWC bf16 [N,H,W,C] (a Linear passes [1,1,M,K]); returns [N,OH,OW,n_out].

    Output selection: bf16 by default, fp32 with `out_f32`, both with `dual` (returns the pair
    (fp32, bf16): the fp32 master of a residual stream plus its bf16 tensor-core operand copy).
    `out` may pre-allocate the destination (a tensor, or an (fp32, bf16) pair for `dual`).
    `stats`: also return the per-32-row-slab (sum, sumsq) of every output column ([M/32, n_out, 2]
    fp32) for `groupnorm(..., stats1=...)`; returned as the last element of the result tuple, or
    None when the pixel grid does not tile cleanly (`conv_stats_supported`).
    `up2`: nearest x2 upsample folded into the conv (taps = 4, `w_packed` = the four parity matrices of
    `pack_up2_weight`, output grid 2H x 2W).  `stride2`: stride-2 conv read through a strided tensor map (output grid
    H/2 x W/2; `pad_lo` 1 = padding 1 all round, 0 = bottom/right only).  `w2`: `a2` is not a concat source but an
    injected tensor on the output grid with its own 1x1 weights (the fused zero-conv injection)."""
    N, H, W, Cc = a.shape
    if a.dtype != BF16 or not _is_nhwc_slice(a):
        raise TypeError("conv_gemm: A must be bf16 NHWC, dense or a channel slice of a dense NHWC tensor")
    if stride2 and (H % 2 or W % 2):
        raise ValueError("conv_gemm: stride2 needs even H and W")
    gh, gw = (H // 2, W // 2) if stride2 else (H, W)                 # the M grid the kernel tiles
    OH, OW = (2 * H, 2 * W) if up2 else (gh, gw)                       # the output grid
    if a2 is not None and (a2.dtype != BF16 or not _is_nhwc_slice(a2) or
                           tuple(a2.shape[:3]) != ((N, OH, OW) if w2 is not None else (N, H, W))):
        raise TypeError("conv_gemm: a2 must be bf16 NHWC with the pixel grid of A (of the output when it is injected)")
    n_cols = n_out // 2 if act == 2 else n_out
    of = oh = None
    if dual:
        of, oh = out if out is not None else (torch.empty((N, OH, OW, n_cols), dtype=torch.float32, device=a.device),
                                              torch.empty((N, OH, OW, n_cols), dtype=BF16, device=a.device))
    elif out is not None:
        of, oh = (out, None) if out.dtype == torch.float32 else (None, out)
    elif out_f32:
        of = torch.empty((N, OH, OW, n_cols), dtype=torch.float32, device=a.device)
    else:
        oh = torch.empty((N, OH, OW, n_cols), dtype=BF16, device=a.device)
    p = ConvParams()
    p.a, p.a_n, p.a_h, p.a_w, p.a_c = _ptr(a), N, gh, gw, Cc
    p.a2, p.a2_c = (_ptr(a2), a2.shape[-1]) if a2 is not None else (None, 0)
    p.a_ld = a.stride(-2)
    p.a2_ld = a2.stride(-2) if a2 is not None else 0
    p.taps = taps
    p.w = _ptr(w_packed)
    p.w_batch_stride = w_batch_stride
    p.w_k, p.w_ld = w_k, w_ld
    p.n_out = n_out
    p.bias = _ptr(bias)
    if row_bias is not None:
        p.row_bias, p.row_bias_ld = _ptr(row_bias), row_bias.stride(0)
    if resid is not None:
        p.resid, p.resid_is_f32, p.ld_resid = _ptr(resid), int(resid.dtype == torch.float32), resid.stride(-2)
    p.alpha = alpha
    p.act = act
    p.act_param = act_param
    p.out_f32, p.out_bf16 = _ptr(of), _ptr(oh)
    ref = of if of is not None else oh
    if of is not None and oh is not None and of.stride(-2) != oh.stride(-2):
        raise ValueError("conv_gemm: dual outputs must share the row stride")
    p.ldo = ref.stride(-2)
    p.tile_n_hint = tile_n
    p.up2 = int(up2)
    p.in_stride2, p.pad_lo = int(stride2), pad_lo
    if w2 is not None:
        if a2 is None:
            raise ValueError("conv_gemm: w2 needs a2")
        p.a2_center, p.w2 = 1, _ptr(w2)
    st = None
    cb1, cb2 = (Cc + 63) // 64, ((a2.shape[-1] + 63) // 64 if a2 is not None else 0)
    kb = taps * cb1 + cb2 if w2 is not None else taps * (cb1 + cb2)
    if stats and act != 2 and (w_batch_stride == 0 or up2) and conv_stats_supported(N, gh, gw, n_out, kb) and \
            (not up2 or gw % 32 == 0):
        st = torch.empty(((N * OH * OW + 31) // 32, n_out, 2), dtype=torch.float32, device=a.device)
        p.stats_out = _ptr(st)
    if split_k:
        ws = _splitk_workspace(a.device)
        p.workspace, p.workspace_bytes = _ptr(ws), ws.numel()
    if GEMM_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.load().rdeic_conv_gemm(C.byref(p), _stream()), "rdeic_conv_gemm")
        e1.record()
        # algorithmic FLOPs of the reference op: an upsample + 3x3 conv counts its 9 taps on the 2H x 2W
        # grid although the folded form executes 4 taps per output pixel
        if up2:
            fl = 2.0 * N * OH * OW * n_out * 9 * Cc
        elif w2 is not None:
            fl = 2.0 * N * gh * gw * n_out * (taps * Cc + a2.shape[-1])
        else:
            fl = 2.0 * N * gh * gw * n_out * taps * (Cc + (a2.shape[-1] if a2 is not None else 0))
        GEMM_PROFILE.append((e0, e1, fl))
    else:
        check(_lib.load().rdeic_conv_gemm(C.byref(p), _stream()), "rdeic_conv_gemm")
    if stats:
        return (of, oh, st) if dual else (ref, st)
    if dual:
        return of, oh
    return ref


def pack_up2_weight(w: torch.Tensor) -> torch.Tensor:
    """conv3x3 weights OIHW fp32 that follow a nearest x2 upsample (openaimodel.py:106-113, model.py:63-67) ->
    bf16 [4, n_out, 4 * cpad]: one 2x2 kernel per output parity class (py, px).  Output pixel (2y+py, 2x+px) reads
    upsampled rows 2y+py-1 .. 2y+py+1, i.e. input rows y-1, y, y (py = 0) or y, y, y+1 (py = 1): taps that land on
    the same input pixel are summed in fp32 before the bf16 rounding.  The padding survives unchanged: the only
    out-of-range input row is y-1 = -1 (py = 0) or y+1 = H (py = 1), exactly the zero rows of the padded 2H grid."""
    w = _need(w, torch.float32, "pack_up2_weight")
    if w.dim() != 4 or tuple(w.shape[2:]) != (3, 3):
        raise ValueError("pack_up2_weight: expected OIHW weights of a 3x3 kernel")

    def fold(t: torch.Tensor, axis: int, par: int) -> torch.Tensor:
        a, b, c = t.unbind(axis)
        return torch.stack([a, b + c], axis) if par == 0 else torch.stack([a + b, c], axis)

    mats = [pack_conv_weight(fold(fold(w, 2, py), 3, px).contiguous()) for py in (0, 1) for px in (0, 1)]
    return torch.stack(mats).contiguous()


_stats_ok = {}


def conv_stats_supported(N: int, H: int, W: int, n_out: int = 32, k_blocks: int = 1) -> bool:
    """Whether conv_gemm(stats=True) will emit GroupNorm statistics for this problem shape."""
    key = (N, H, W, n_out, k_blocks)
    if key not in _stats_ok:
        _stats_ok[key] = bool(_lib.load().rdeic_conv_stats_supported(N, H, W, n_out, k_blocks)) and (H * W) % 32 == 0
    return _stats_ok[key]


def _is_nhwc_slice(t: torch.Tensor) -> bool:
    """dense [N,H,W,C], or channels [c0, c0+C) of a dense [N,H,W,ld] tensor."""
    N, H, W, _ = t.shape
    ld = t.stride(-2)
    return t.stride(-1) == 1 and ld >= t.shape[-1] and (H == 1 or t.stride(1) == W * ld) and (N == 1 or t.stride(0) == H * W * ld)


def linear(x: torch.Tensor, w_packed: torch.Tensor, n_out: int, **kw):
    """x [..., K] bf16 -> [..., n_out] through the same tensor-core kernel (taps = 1)."""
    K = x.shape[-1]
    M = x.numel() // K
    resid = kw.pop("resid", None)
    if resid is not None:
        resid = resid.reshape(1, 1, M, resid.shape[-1])
    out = kw.pop("out", None)
    if out is not None:
        out = tuple(o.view(1, 1, M, o.shape[-1]) for o in out) if isinstance(out, tuple) else out.view(1, 1, M, out.shape[-1])
    stats = kw.get("stats", False)
    y = conv_gemm(x.reshape(1, 1, M, K), w_packed, n_out, 1, resid=resid, out=out, **kw)
    st = None
    if stats:
        *y, st = y
        y = y[0] if len(y) == 1 else tuple(y)
    if isinstance(y, tuple):
        y = tuple(t.view(*x.shape[:-1], t.shape[-1]) for t in y)
        return y + (st,) if stats else y
    y = y.view(*x.shape[:-1], y.shape[-1])
    return (y, st) if stats else y


def attention(q, k, v, heads: int, d: int, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B,Nq,>=heads*d], k/v [B,Nk,...] bf16 (may be column slices of a fused projection)."""
    B, Nq = q.shape[0], q.shape[1]
    Nk = k.shape[1]
    if out is None:
        out = torch.empty((B, Nq, heads * d), dtype=BF16, device=q.device)
    for name, t in (("q", q), ("k", k), ("v", v)):
        if t.dtype != BF16 or t.stride(-1) != 1:
            raise TypeError(f"attention: {name} must be bf16 with unit inner stride")
    check(_lib.load().rdeic_attention(_ptr(q), _ptr(k), _ptr(v), _ptr(out), B, heads, Nq, Nk, d, q.stride(1),
                                      k.stride(1), v.stride(1), out.stride(1), q.stride(0), k.stride(0), v.stride(0),
                                      out.stride(0), scale, _stream()), "rdeic_attention")
    return out
