"""IDX tensor files (the MNIST container): 2 zero bytes, a dtype code, the rank, big-endian uint32 dims, raw big-endian data.

Reference: helpers/idx.py:14-95 (same six dtype codes, `npz2idxtar` packing an .npz as a tar of idx members). Files ending
in .gz are handled transparently, which the reference leaves to `gunzip`.
"""
from __future__ import annotations

import gzip
import io
import struct
import tarfile
from typing import BinaryIO

import numpy as np

_CODES = {0x08: np.uint8, 0x09: np.int8, 0x0B: np.int16, 0x0C: np.int32, 0x0D: np.float32, 0x0E: np.float64}
_BY_TYPE = {np.dtype(t): c for c, t in _CODES.items()}


def _code_of(dtype) -> int:
    try:
        return _BY_TYPE[np.dtype(dtype)]
    except KeyError:
        raise ValueError(f"idx cannot store dtype {dtype}") from None


def _type_of(code: int):
    try:
        return _CODES[code]
    except KeyError:
        raise ValueError(f"unknown idx dtype code 0x{code:02x}") from None


def write_idx_header(f: BinaryIO, a: np.ndarray) -> None:
    if a.ndim > 255:
        raise ValueError("idx rank is one byte")
    f.write(struct.pack(">BBBB", 0, 0, _code_of(a.dtype), a.ndim))
    f.write(struct.pack(f">{a.ndim}I", *a.shape))


def write_idx_to(f: BinaryIO, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a)
    write_idx_header(f, a)
    f.write(a.astype(a.dtype.newbyteorder(">"), copy=False).tobytes())


def write_idx_file(name: str, a: np.ndarray) -> None:
    opener = gzip.open if str(name).endswith(".gz") else open
    with opener(name, "wb") as f:
        write_idx_to(f, a)


def read_idx_header(f: BinaryIO):
    head = f.read(4)
    if len(head) != 4 or head[0] != 0 or head[1] != 0:
        raise ValueError("not an idx stream (bad magic)")
    dtype, rank = _type_of(head[2]), head[3]
    dims = struct.unpack(f">{rank}I", f.read(4 * rank)) if rank else ()
    return np.dtype(dtype), tuple(dims)


def read_idx_from(f: BinaryIO) -> np.ndarray:
    dtype, dims = read_idx_header(f)
    count = int(np.prod(dims, dtype=np.int64)) if dims else 1
    raw = f.read(count * dtype.itemsize)
    if len(raw) != count * dtype.itemsize:
        raise ValueError("truncated idx payload")
    return np.frombuffer(raw, dtype=dtype.newbyteorder(">")).astype(dtype).reshape(dims)


def read_idx_file(name: str) -> np.ndarray:
    opener = gzip.open if str(name).endswith(".gz") else open
    with opener(name, "rb") as f:
        return read_idx_from(f)


def npz2idxtar(in_file: str, out_file: str | None = None) -> str:
    """Repack every array of an .npz as an idx member of a tar (checkpoint exchange with non-Python tools)."""
    if out_file is None:
        base = str(in_file)
        out_file = (base[:-4] if base.endswith(".npz") else base) + ".idx.tar"
    with np.load(in_file) as ws, tarfile.open(out_file, "w") as tar:
        for key in ws.files:
            blob = io.BytesIO()
            write_idx_to(blob, ws[key])
            member = tarfile.TarInfo(key)
            member.size = blob.tell()
            blob.seek(0)
            tar.addfile(member, blob)
    return out_file


def idxtar2dict(in_file: str) -> dict:
    """Inverse of `npz2idxtar`."""
    out = {}
    with tarfile.open(in_file, "r") as tar:
        for member in tar.getmembers():
            out[member.name] = read_idx_from(tar.extractfile(member))
    return out
