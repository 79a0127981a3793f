"""Bitstream container of the reference (utils/utils.py:24-74): big-endian `uint32 h, w, n_strings`
followed by length-prefixed byte strings.  Byte-for-byte the reference's file format, so streams
written by either side are read by the other."""
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
