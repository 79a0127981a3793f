"""Oracle: entropy-model front end (numpy, bit-exact integer/byte work).  TEST INFRASTRUCTURE.

Follows utils/ckbd.py:6-73 (checkerboard ops), utils/func.py:10-13 (scale table),
compressai 1.2.4 `GaussianConditional.build_indexes` / `EntropyModel.quantize` as called from
utils/ckbd.py:81-82,92-93,102,111 (parity unpinned for compressai: restated from its published
source), model/compression_modules.py:309-338 (VectorQuantiser.quant / get_codebook_entry).
"""
from __future__ import annotations

import math

import numpy as np


def ckbd_anchor(y: np.ndarray) -> np.ndarray:
    """utils/ckbd.py:35-39."""
    out = np.zeros_like(y)
    out[:, :, 0::2, 1::2] = y[:, :, 0::2, 1::2]
    out[:, :, 1::2, 0::2] = y[:, :, 1::2, 0::2]
    return out


def ckbd_nonanchor(y: np.ndarray) -> np.ndarray:
    """utils/ckbd.py:41-45."""
    out = np.zeros_like(y)
    out[:, :, 0::2, 0::2] = y[:, :, 0::2, 0::2]
    out[:, :, 1::2, 1::2] = y[:, :, 1::2, 1::2]
    return out


def ckbd_split(y):
    """utils/ckbd.py:6-24."""
    return ckbd_anchor(y), ckbd_nonanchor(y)


def ckbd_merge(anchor, nonanchor):
    """utils/ckbd.py:26-33."""
    return anchor + nonanchor


def ckbd_anchor_sequeeze(y):
    """utils/ckbd.py:47-52."""
    B, C, H, W = y.shape
    out = np.zeros((B, C, H, W // 2), dtype=np.float32)
    out[:, :, 0::2, :] = y[:, :, 0::2, 1::2]
    out[:, :, 1::2, :] = y[:, :, 1::2, 0::2]
    return out


def ckbd_nonanchor_sequeeze(y):
    """utils/ckbd.py:54-59."""
    B, C, H, W = y.shape
    out = np.zeros((B, C, H, W // 2), dtype=np.float32)
    out[:, :, 0::2, :] = y[:, :, 0::2, 0::2]
    out[:, :, 1::2, :] = y[:, :, 1::2, 1::2]
    return out


def ckbd_anchor_unsequeeze(a):
    """utils/ckbd.py:61-66."""
    B, C, H, W = a.shape
    out = np.zeros((B, C, H, W * 2), dtype=np.float32)
    out[:, :, 0::2, 1::2] = a[:, :, 0::2, :]
    out[:, :, 1::2, 0::2] = a[:, :, 1::2, :]
    return out


def ckbd_nonanchor_unsequeeze(n):
    """utils/ckbd.py:68-73."""
    B, C, H, W = n.shape
    out = np.zeros((B, C, H, W * 2), dtype=np.float32)
    out[:, :, 0::2, 0::2] = n[:, :, 0::2, :]
    out[:, :, 1::2, 1::2] = n[:, :, 1::2, :]
    return out


def get_scale_table(lo: float = 0.11, hi: float = 256.0, levels: int = 64) -> np.ndarray:
    """utils/func.py:10-13: torch.exp(torch.linspace(log(min), log(max), levels)) in fp32.
    Computed with torch so the fp32 table is bit-identical to the reference's."""
    import torch

    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels)).numpy()


def build_indexes(scales: np.ndarray, table: np.ndarray, lower_bound: float = 0.11) -> np.ndarray:
    """compressai 1.2.4 GaussianConditional.build_indexes: scales = LowerBound(0.11)(scales);
    indexes = full(len(table)-1); for s in table[:-1]: indexes -= (scales <= s)."""
    s = np.where(np.isnan(scales), scales, np.maximum(scales, np.float32(lower_bound))).astype(np.float32)
    idx = np.full(s.shape, len(table) - 1, dtype=np.int32)
    for t in table[:-1]:
        idx -= (s <= np.float32(t)).astype(np.int32)
    return idx


def quantize_symbols(x: np.ndarray, means: np.ndarray | None) -> np.ndarray:
    """compressai 1.2.4 EntropyModel.quantize(x, "symbols", means): (x - means).round().int();
    torch.round is round-half-to-even == np.rint."""
    v = x.astype(np.float32)
    if means is not None:
        v = (v - means.astype(np.float32)).astype(np.float32)
    return np.rint(v).astype(np.int32)


def dequantize(symbols: np.ndarray, means: np.ndarray) -> np.ndarray:
    """utils/ckbd.py:85,96,104,113: symbols.float() + means."""
    return (symbols.astype(np.float32) + means.astype(np.float32)).astype(np.float32)


def compress_phase(y, scales, means, table, which: int, lower_bound: float = 0.11):
    """utils/ckbd.py:76-97 without the rANS call: returns (symbols, indexes, y_hat)."""
    sq = ckbd_anchor_sequeeze if which == 0 else ckbd_nonanchor_sequeeze
    unsq = ckbd_anchor_unsequeeze if which == 0 else ckbd_nonanchor_unsequeeze
    y_s, s_s, m_s = sq(y), sq(scales), sq(means)
    idx = build_indexes(s_s, table, lower_bound)
    sym = quantize_symbols(y_s, m_s)
    return sym, idx, unsq(dequantize(sym, m_s))


def decompress_phase_pre(scales, means, table, which: int, lower_bound: float = 0.11):
    """utils/ckbd.py:99-102,108-111: (means_squeezed, indexes)."""
    sq = ckbd_anchor_sequeeze if which == 0 else ckbd_nonanchor_sequeeze
    return sq(means), build_indexes(sq(scales), table, lower_bound)


def decompress_phase_post(symbols, means_sq, which: int):
    """utils/ckbd.py:104-105,113-114."""
    unsq = ckbd_anchor_unsequeeze if which == 0 else ckbd_nonanchor_unsequeeze
    return unsq(dequantize(symbols, means_sq))


def vq_quant(z: np.ndarray, codebook: np.ndarray):
    """model/compression_modules.py:309-331: argmin_j |z|^2 + |e_j|^2 - 2 z.e_j (first min)."""
    B, D, H, W = z.shape
    zf = np.ascontiguousarray(z.transpose(0, 2, 3, 1)).reshape(-1, D).astype(np.float32)
    e = codebook.astype(np.float32)
    d = (zf ** 2).sum(1, keepdims=True) + (e ** 2).sum(1)[None, :] - 2.0 * (zf @ e.T)
    idx = np.argmin(d, axis=1).astype(np.int64)
    zq = e[idx].reshape(B, H, W, D).transpose(0, 3, 1, 2)
    return np.ascontiguousarray(zq), idx.reshape(B, H, W)


def vq_lookup(idx: np.ndarray, codebook: np.ndarray) -> np.ndarray:
    """model/compression_modules.py:333-338."""
    B, H, W = idx.shape
    return np.ascontiguousarray(codebook[idx.reshape(-1)].reshape(B, H, W, -1).transpose(0, 3, 1, 2))
