"""Drop-in for the reference's `utils/ckbd.py`: checkerboard split / merge / squeeze / unsqueeze and
the quantise -> symbols / CDF-index glue around the (host) rANS coder, on CUDA tensors, bit-exact.

Function names keep the reference's spelling (`sequeeze`).  The four `(de)compress_*` helpers keep
the reference signatures (utils/ckbd.py:76-115) but run each phase as ONE fused kernel and hand
symbols / indexes to the coder as int32 numpy views of pinned host buffers instead of
`.reshape(-1).tolist()` Python lists built from a synchronous D2H copy per tensor (ckbd.py:83-84,103).
The byte coder itself (compressai rANS, torchac) stays in the host library: out of scope.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import ops

ANCHOR, NONANCHOR = 0, 1


def ckbd_split(y):
    """utils/ckbd.py:6-24."""
    return ops.ckbd_split(y)


def ckbd_merge(anchor, nonanchor):
    """utils/ckbd.py:26-33."""
    return ops.ckbd_merge(anchor, nonanchor)


def ckbd_anchor(y):
    """utils/ckbd.py:35-39."""
    return ops.ckbd_mask(y, ANCHOR)


def ckbd_nonanchor(y):
    """utils/ckbd.py:41-45."""
    return ops.ckbd_mask(y, NONANCHOR)


def ckbd_anchor_sequeeze(y):
    """utils/ckbd.py:47-52."""
    return ops.ckbd_squeeze(y, ANCHOR)


def ckbd_nonanchor_sequeeze(y):
    """utils/ckbd.py:54-59."""
    return ops.ckbd_squeeze(y, NONANCHOR)


def ckbd_anchor_unsequeeze(anchor):
    """utils/ckbd.py:61-66."""
    return ops.ckbd_unsqueeze(anchor, ANCHOR)


def ckbd_nonanchor_unsequeeze(nonanchor):
    """utils/ckbd.py:68-73."""
    return ops.ckbd_unsqueeze(nonanchor, NONANCHOR)


def get_scale_table(min=0.11, max=256, levels=64):  # noqa: A002  (reference argument names)
    """utils/func.py:10-13 — evaluated by torch on the host so the fp32 table is the reference's."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class GaussianConditional:
    """The two arithmetic entry points of compressai 1.2.4 `GaussianConditional` the decode path uses
    (`build_indexes`, `quantize`), over CUDA tensors.  `scale_bound` is compressai's LowerBound(0.11)."""

    def __init__(self, scale_table: Optional[torch.Tensor] = None, scale_bound: float = 0.11, device="cuda"):
        table = get_scale_table() if scale_table is None else torch.as_tensor(scale_table, dtype=torch.float32)
        self.scale_table = table.to(device, torch.float32).contiguous()
        self.scale_bound = float(scale_bound)

    def update_scale_table(self, scale_table, force=False):
        """compressai GaussianConditional.update_scale_table: new table + the rANS coder's quantised CDF
        tables.  Those tables are consumed only by compressai's own coder, so they are built by
        compressai when it is installed (`cdf_tables()`); the index / symbol arithmetic needs only
        the scale table."""
        self.scale_table = torch.as_tensor(scale_table, dtype=torch.float32).to(self.scale_table.device).contiguous()
        self._tables = None
        return True

    def cdf_tables(self):
        """(quantized_cdf, cdf_lengths, offsets) as the Python lists compression.py:163-165 hands to the
        rANS coder, or (None, None, None) without compressai (loopback coders ignore them)."""
        if getattr(self, "_tables", None) is None:
            try:
                from compressai.entropy_models import GaussianConditional as _GC
            except ImportError:
                return None, None, None
            g = _GC(None)
            g.update_scale_table(self.scale_table.cpu(), force=True)
            self._tables = (g.quantized_cdf.tolist(), g.cdf_length.reshape(-1).int().tolist(),
                            g.offset.reshape(-1).int().tolist())
        return self._tables

    def build_indexes(self, scales: torch.Tensor) -> torch.Tensor:
        return ops.build_indexes(scales, self.scale_table, self.scale_bound)

    def quantize(self, inputs: torch.Tensor, mode: str, means: Optional[torch.Tensor] = None) -> torch.Tensor:
        if mode == "symbols":
            return ops.quantize_symbols(inputs, means)
        if mode == "dequantize":
            sym = ops.quantize_symbols(inputs, means)
            zeros = means if means is not None else torch.zeros_like(inputs)
            return ops.dequantize(sym, zeros.expand_as(inputs).contiguous())
        raise ValueError(f'Invalid quantization mode: "{mode}"')

    def dequantize(self, inputs: torch.Tensor, means: Optional[torch.Tensor] = None) -> torch.Tensor:
        means = torch.zeros(inputs.shape, dtype=torch.float32, device=inputs.device) if means is None else means
        return ops.dequantize(inputs.to(torch.int32), means.expand(inputs.shape).contiguous())


class TorchacHyperLatentCoder:
    """utils/ckbd.py:118-141: the VQ indices of the hyper latent are coded by torchac against a
    uniform CDF over the codebook.  torchac is a host library (out of scope, like the rANS coder);
    this wrapper only restates the reference's calls into it."""

    def __init__(self, codebook_size: int):
        self.codebook_size = codebook_size

    def _cdf(self, shape):
        b, h, w = shape
        k = self.codebook_size
        cdf = torch.cumsum(torch.full((k,), 1.0 / k), dim=0)
        cdf = torch.cat([torch.zeros(1), cdf]).view(1, 1, 1, -1).expand(b, h, w, -1).clone()
        cdf[..., -1] = 1.0
        return cdf

    @staticmethod
    def _torchac():
        try:
            import torchac
        except ImportError as e:
            raise RuntimeError("the hyper-latent byte coder needs torchac (utils/ckbd.py:4); install it or pass "
                               "hyper_latent_coder= to Compression") from e
        return torchac

    def compress(self, encoding_indices: torch.Tensor):
        ac = self._torchac()
        return ac.encode_float_cdf(self._cdf(encoding_indices.shape), encoding_indices.to(torch.int16).to("cpu"),
                                   check_input_bounds=True)

    def decompress(self, string, shape) -> torch.Tensor:
        h, w = shape
        return self._torchac().decode_float_cdf(self._cdf((1, int(h), int(w))), string)


class _Pinned:
    """Reusable pinned int32 staging buffers for the symbol hand-off."""

    def __init__(self):
        self.buf = {}

    def get(self, n: int, slot: str) -> torch.Tensor:
        b = self.buf.get(slot)
        if b is None or b.numel() < n:
            b = torch.empty(max(n, 1 << 16), dtype=torch.int32).pin_memory()
            self.buf[slot] = b
        return b[:n]

    def to_host(self, t: torch.Tensor, slot: str) -> torch.Tensor:
        b = self.get(t.numel(), slot)
        b.copy_(t.reshape(-1), non_blocking=True)
        return b

    def sync(self):
        torch.cuda.current_stream().synchronize()


_pinned = _Pinned()


class SymbolStream:
    """GPU-resident stand-in for the `symbols_list` / `indexes_list` Python lists of
    model/compression.py:167-168.  The encode-phase kernels write their int32
    output straight into consecutive windows of one device buffer in stream order; nothing crosses
    PCIe and nothing synchronises until `host()` — ONE pinned copy for the whole image batch instead
    of 20 `.reshape(-1).tolist()` round trips (utils/ckbd.py:83-84,94-95)."""

    def __init__(self, capacity: int, device):
        self.buf = torch.empty(capacity, dtype=torch.int32, device=device)
        self.n = 0

    def window(self, n: int) -> torch.Tensor:
        if self.n + n > self.buf.numel():
            raise RuntimeError("SymbolStream: capacity exceeded")
        w = self.buf[self.n:self.n + n]
        self.n += n
        return w

    def host(self, slot: str) -> np.ndarray:
        """int32 numpy view of a pinned buffer (valid until the next `host()` of the same slot)."""
        h = _pinned.to_host(self.buf[:self.n], slot)
        _pinned.sync()
        return h.numpy()


def _phase(gc: GaussianConditional, y, scales, means, symbols_list, indexes_list, which: int):
    if isinstance(symbols_list, SymbolStream):
        n = y.numel() // 2
        _, _, y_hat = ops.ckbd_encode_phase(y, scales, means, gc.scale_table, gc.scale_bound, which,
                                            sym_out=symbols_list.window(n), idx_out=indexes_list.window(n))
        return y_hat
    sym, idx, y_hat = ops.ckbd_encode_phase(y, scales, means, gc.scale_table, gc.scale_bound, which)
    hs, hi = _pinned.to_host(sym, "sym"), _pinned.to_host(idx, "idx")
    _pinned.sync()
    symbols_list.extend(hs.tolist())
    indexes_list.extend(hi.tolist())
    return y_hat


def compress_anchor(gaussian_conditional, anchor, scales_anchor, means_anchor, symbols_list, indexes_list):
    """utils/ckbd.py:76-86."""
    return _phase(gaussian_conditional, anchor, scales_anchor, means_anchor, symbols_list, indexes_list, ANCHOR)


def compress_nonanchor(gaussian_conditional, nonanchor, scales_nonanchor, means_nonanchor, symbols_list, indexes_list):
    """utils/ckbd.py:88-97."""
    return _phase(gaussian_conditional, nonanchor, scales_nonanchor, means_nonanchor, symbols_list, indexes_list,
                  NONANCHOR)


def _dephase(gc: GaussianConditional, scales, means, decoder, cdf, cdf_lengths, offsets, which: int):
    """One host hand-off per phase is inherent (the coder needs this phase's indexes to produce the
    symbols the next GPU phase depends on); it goes through pinned int32 buffers both ways.  A coder
    that sets `accepts_arrays = True` receives / returns int32 numpy arrays (views of those pinned
    buffers); compressai's pybind coder gets the Python list it expects."""
    means_sq, idx = ops.ckbd_squeeze_indexes(scales, means, gc.scale_table, gc.scale_bound, which)
    hi = _pinned.to_host(idx, "idx")
    _pinned.sync()
    arrays = getattr(decoder, "accepts_arrays", False)
    symbols = decoder.decode_stream(hi.numpy() if arrays else hi.tolist(), cdf, cdf_lengths, offsets)
    hs = _pinned.get(idx.numel(), "sym_in")
    hs.numpy()[:] = np.asarray(symbols, dtype=np.int32)
    sym = hs.to(means_sq.device, non_blocking=True).view(means_sq.shape)
    return ops.ckbd_decode_phase(sym, means_sq, which)


def decompress_anchor(gaussian_conditional, scales_anchor, means_anchor, decoder, cdf, cdf_lengths, offsets):
    """utils/ckbd.py:99-106."""
    return _dephase(gaussian_conditional, scales_anchor, means_anchor, decoder, cdf, cdf_lengths, offsets, ANCHOR)


def decompress_nonanchor(gaussian_conditional, scales_nonanchor, means_nonanchor, decoder, cdf, cdf_lengths, offsets):
    """utils/ckbd.py:108-115."""
    return _dephase(gaussian_conditional, scales_nonanchor, means_nonanchor, decoder, cdf, cdf_lengths, offsets,
                    NONANCHOR)
