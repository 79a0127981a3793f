"""Bitstream container of the reference (utils/utils.py:24-74): big-endian `uint32 h, w, n_strings`
followed by length-prefixed byte strings.  The container framing is byte-for-byte the reference's.  The
entropy-coded payload inside decodes wherever the entropy-parameter nets reproduce the encoder's CDF indexes:
`Compression(precision="mixed")` (the default) and `"fp32"` run those nets in fp32 like the reference, so
streams are exchangeable with it; a stream written with `precision="bf16"` is only decodable by that mode
(determinism contract)."""
from __future__ import annotations

import struct
from pathlib import Path
from typing import BinaryIO, List, Sequence, Tuple


def write_uints(fd: BinaryIO, values: Sequence[int], fmt: str = ">{:d}I") -> int:
    fd.write(struct.pack(fmt.format(len(values)), *values))
    return len(values) * 4


def read_uints(fd: BinaryIO, n: int, fmt: str = ">{:d}I") -> Tuple[int, ...]:
    data = fd.read(n * 4)
    if len(data) != n * 4:
        raise struct.error(f"unpack requires a buffer of {n * 4} bytes")
    return struct.unpack(fmt.format(n), data)


def write_bytes(fd: BinaryIO, values: bytes, fmt: str = ">{:d}s"):
    if len(values) == 0:
        return None
    fd.write(struct.pack(fmt.format(len(values)), values))
    return len(values)


def read_bytes(fd: BinaryIO, n: int, fmt: str = ">{:d}s") -> bytes:
    data = fd.read(n)
    if len(data) != n:
        raise struct.error(f"unpack requires a buffer of {n} bytes")
    return struct.unpack(fmt.format(n), data)[0]


def read_body(fd: BinaryIO):
    """utils/utils.py:57-65 -> (lstrings, shape): lstrings = [[s0], [s1], ...]."""
    lstrings: List[List[bytes]] = []
    shape = read_uints(fd, 2)
    n_strings = read_uints(fd, 1)[0]
    for _ in range(n_strings):
        lstrings.append([read_bytes(fd, read_uints(fd, 1)[0])])
    return lstrings, shape


def write_body(fd: BinaryIO, shape, out_strings) -> int:
    """utils/utils.py:68-74 -> bytes written (a zero-length string adds only its length word; the
    reference then fails on `+= None`, here it is counted as 0)."""
    bytes_cnt = write_uints(fd, (int(shape[0]), int(shape[1]), len(out_strings)))
    for s in out_strings:
        bytes_cnt += write_uints(fd, (len(s[0]),))
        bytes_cnt += write_bytes(fd, s[0]) or 0
    return bytes_cnt


def filesize(filepath: str) -> int:
    """utils/utils.py:77-80."""
    if not Path(filepath).is_file():
        raise ValueError(f'Invalid file "{filepath}".')
    return Path(filepath).stat().st_size


def pad(img, scale: int):
    """utils/image/common.py:251-258: zero-pad an HWC numpy image at the bottom / right to multiples of `scale`."""
    import math

    import numpy as np

    h, w = img.shape[:2]
    ph = 0 if h % scale == 0 else math.ceil(h / scale) * scale - h
    pw = 0 if w % scale == 0 else math.ceil(w / scale) * scale - w
    return np.pad(img, pad_width=((0, ph), (0, pw), (0, 0)), mode="constant", constant_values=0)


def group_by_padded_size(sizes, batch_size: int, scale: int = 64):
    """inference_partition.py:438-452: bucket images by their padded (H, W) and cut each bucket into
    batches, in the reference's order (buckets sorted by key, members in input order).
    sizes: [(h, w), ...] -> [((pad_h, pad_w), [indices...]), ...]."""
    groups = {}
    for i, (h, w) in enumerate(sizes):
        key = ((h + scale - 1) // scale * scale, (w + scale - 1) // scale * scale)
        groups.setdefault(key, []).append(i)
    bs = max(1, int(batch_size))
    return [(key, groups[key][i:i + bs]) for key in sorted(groups) for i in range(0, len(groups[key]), bs)]
