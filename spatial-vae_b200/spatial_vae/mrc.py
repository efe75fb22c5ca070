"""Minimal MRC / MRCS stack reader and writer (host IO; reference spatial_vae/mrc.py:7-218).

Only what the particle CLI needs: the 1024-byte header (first 10 int32 words, the extended-header
length at byte 92, the data mode), the extended header skipped, the voxels as a numpy array.
"""
import struct
from collections import namedtuple

import numpy as np

MODES = {0: np.int8, 1: np.int16, 2: np.float32, 4: np.complex64, 6: np.uint16}
MRCHeader = namedtuple('MRCHeader', 'nx ny nz mode next amin amax amean')


def parse(content):
    """bytes -> (array (nz, ny, nx) or (ny, nx) when nz == 1, header, extended header bytes)."""
    nx, ny, nz, mode = struct.unpack_from('<4i', content, 0)
    amin, amax, amean = struct.unpack_from('<3f', content, 76)
    ext = struct.unpack_from('<i', content, 92)[0]
    if mode not in MODES:
        raise ValueError(f'unsupported MRC mode {mode}')
    header = MRCHeader(nx, ny, nz, mode, ext, amin, amax, amean)
    start = 1024 + ext
    array = np.frombuffer(content, dtype=MODES[mode], offset=start, count=nx * ny * nz).reshape(nz, ny, nx)
    return (array[0] if nz == 1 else array), header, content[1024:start]


def write(f, array):
    """Write a float32 stack with a bare header (mode 2)."""
    array = np.asarray(array, dtype=np.float32)
    if array.ndim == 2:
        array = array[None]
    nz, ny, nx = array.shape
    head = bytearray(1024)
    struct.pack_into('<4i', head, 0, nx, ny, nz, 2)
    struct.pack_into('<3i', head, 28, nx, ny, nz)
    struct.pack_into('<3f', head, 40, float(nx), float(ny), float(nz))
    struct.pack_into('<3f', head, 52, 90.0, 90.0, 90.0)
    struct.pack_into('<3i', head, 64, 1, 2, 3)
    struct.pack_into('<3f', head, 76, float(array.min()), float(array.max()), float(array.mean()))
    head[208:212] = b'MAP '
    head[212:216] = bytes([0x44, 0x44, 0, 0])
    f.write(bytes(head))
    f.write(array.tobytes())
