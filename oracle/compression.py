"""Oracle: slice orchestration of the entropy front end (numpy/torch CPU).  TEST INFRASTRUCTURE.

Follows model/compression.py:161-206 (compress) and :233-266 (decompress) with the learned conv
stacks abstracted as callables and the rANS coder replaced by a loopback."""
from __future__ import annotations

import numpy as np
import torch

from . import entropy as oe


class LoopbackCoder:
    """Stands in for compressai BufferedRansEncoder / RansDecoder: stores (symbols, indexes) on
    encode and replays the symbols on decode after checking the decoder presents the SAME indexes
    (which is what keeps a real arithmetic coder in sync)."""

    def __init__(self):
        self.symbols, self.indexes, self.pos = [], [], 0

    def encode_with_indexes(self, symbols, indexes, *a):
        self.symbols, self.indexes, self.pos = list(symbols), list(indexes), 0

    def decode_stream(self, indexes, *a):
        n = len(indexes)
        want = self.indexes[self.pos:self.pos + n]
        if list(indexes) != want:
            raise RuntimeError("decoder CDF indexes diverged from the encoder's")
        out = self.symbols[self.pos:self.pos + n]
        self.pos += n
        return out


def _params(idx, hyper, y_hat_slices, fns, slice_anchor=None):
    ep_a, ep_n, local_ctx, channel_ctx = fns
    ctx = [] if idx == 0 else [channel_ctx[idx](torch.cat(y_hat_slices, dim=1))]
    if slice_anchor is None:
        p = ep_a[idx](torch.cat(ctx + [hyper], dim=1) if ctx else hyper)
    else:
        p = ep_n[idx](torch.cat([local_ctx[idx](slice_anchor)] + ctx + [hyper], dim=1))
    s, m = p.chunk(2, 1)
    return s.contiguous(), m.contiguous()


def compress(y, hyper, slice_ch, fns, table):
    symbols, indexes, y_hat_slices, off = [], [], [], 0
    for idx, c in enumerate(slice_ch):
        ys = y[:, off:off + c].contiguous().numpy()
        off += c
        a, n = oe.ckbd_split(ys)
        sa, ma = _params(idx, hyper, y_hat_slices, fns)
        sym, ind, a_hat = oe.compress_phase(a, sa.numpy(), ma.numpy(), table, 0)
        symbols += sym.reshape(-1).tolist(); indexes += ind.reshape(-1).tolist()
        sn, mn = _params(idx, hyper, y_hat_slices, fns, torch.from_numpy(a_hat))
        sym, ind, n_hat = oe.compress_phase(n, sn.numpy(), mn.numpy(), table, 1)
        symbols += sym.reshape(-1).tolist(); indexes += ind.reshape(-1).tolist()
        y_hat_slices.append(torch.from_numpy(oe.ckbd_merge(a_hat, n_hat)))
    return symbols, indexes, torch.cat(y_hat_slices, dim=1)


def decompress(hyper, slice_ch, fns, table, coder):
    y_hat_slices = []
    for idx in range(len(slice_ch)):
        sa, ma = _params(idx, hyper, y_hat_slices, fns)
        msq, ind = oe.decompress_phase_pre(sa.numpy(), ma.numpy(), table, 0)
        sym = np.asarray(coder.decode_stream(ind.reshape(-1).tolist()), dtype=np.int32).reshape(msq.shape)
        a_hat = oe.decompress_phase_post(sym, msq, 0)
        sn, mn = _params(idx, hyper, y_hat_slices, fns, torch.from_numpy(a_hat))
        msq, ind = oe.decompress_phase_pre(sn.numpy(), mn.numpy(), table, 1)
        sym = np.asarray(coder.decode_stream(ind.reshape(-1).tolist()), dtype=np.int32).reshape(msq.shape)
        n_hat = oe.decompress_phase_post(sym, msq, 1)
        y_hat_slices.append(torch.from_numpy(oe.ckbd_merge(n_hat, a_hat)))
    return torch.cat(y_hat_slices, dim=1)
