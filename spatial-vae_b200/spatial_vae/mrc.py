"""MRC / MRCS stack reader and writer (host IO; same functions and argument meaning as reference
spatial_vae/mrc.py: MRCHeader, parse, get_mode, make_header, write).

The 1024-byte header is described by ONE table (field name, struct code) in file order, following the MRC2014 / IMOD
layout the reference uses (mrc.py:7-106); the struct format and the MRCHeader namedtuple are generated from it.
Files written with default arguments are byte-identical to the reference's (pinned by tests/golden/ingest.npz).
"""
import struct
from collections import namedtuple

import numpy as np

# (names, struct code); None = padding.  Offsets in bytes in the comments.
_LAYOUT = [
    ("nx ny nz", "3i"),                                        # 0    columns, rows, sections
    ("mode", "i"),                                             # 12   voxel type
    ("nxstart nystart nzstart", "3i"),                         # 16
    ("mx my mz", "3i"),                                        # 28   grid sampling
    ("xlen ylen zlen", "3f"),                                  # 40   cell size
    ("alpha beta gamma", "3f"),                                # 52   cell angles
    ("mapc mapr maps", "3i"),                                  # 64   axis order
    ("amin amax amean", "3f"),                                 # 76   density statistics
    ("ispg next", "2i"),                                       # 88   space group, extended-header bytes
    ("creatid", "h"),                                          # 96
    (None, "30x"),                                             # 98
    ("nint nreal", "2h"),                                      # 128
    (None, "20x"),                                             # 132
    ("imodStamp imodFlags", "2i"),                             # 152
    ("idtype lens nd1 nd2 vd1 vd2", "6h"),                     # 160
    ("tilt_ox tilt_oy tilt_oz tilt_cx tilt_cy tilt_cz", "6f"),  # 172
    ("xorg yorg zorg", "3f"),                                  # 196
    ("cmap stamp", "4s4s"),                                    # 208
    ("rms", "f"),                                              # 216
    ("nlabl", "i"),                                            # 220
    ("labels", "800s"),                                        # 224 .. 1024
]
header_struct = struct.Struct("<" + "".join(code for _, code in _LAYOUT))
assert header_struct.size == 1024
MRCHeader = namedtuple("MRCHeader", " ".join(names for names, _ in _LAYOUT if names))

# mode <-> voxel dtype: 3 = complex as two int16, 16 = RGB bytes (structured sub-array dtypes)
_MODE_DTYPES = [(0, np.dtype(np.int8)), (1, np.dtype(np.int16)), (2, np.dtype(np.float32)), (3, np.dtype("2h")),
                (4, np.dtype(np.complex64)), (6, np.dtype(np.uint16)), (16, np.dtype("3B"))]


def get_mode(dtype):
    """MRC mode number of a numpy dtype."""
    dtype = np.dtype(dtype)
    for mode, dt in _MODE_DTYPES:
        if dt == dtype:
            return mode
    raise ValueError("MRC incompatible dtype: " + str(dtype))


def parse(content):
    """bytes -> (array (nz, ny, nx), or (ny, nx) when nz == 1; MRCHeader; extended-header bytes)."""
    header = MRCHeader._make(header_struct.unpack(content[:1024]))
    start = 1024 + header.next
    dtypes = dict(_MODE_DTYPES)
    if header.mode not in dtypes:
        raise ValueError(f"unsupported MRC mode {header.mode}")
    array = np.frombuffer(content[start:], dtype=dtypes[header.mode])
    array = array.reshape(header.nz, header.ny, header.nx, *array.shape[1:])
    return (array[0] if header.nz == 1 else array), header, content[1024:start]


def _header(nx, ny, nz, mode, mz, cella, cellb, stats, rms, ispg, exthd_size):
    zero = dict.fromkeys(MRCHeader._fields, 0)
    zero.update(nx=nx, ny=ny, nz=nz, mode=mode, mx=1, my=1, mz=mz, xlen=cella[0], ylen=cella[1], zlen=cella[2],
                alpha=cellb[0], beta=cellb[1], gamma=cellb[2], mapc=1, mapr=2, maps=3, amin=stats[0], amax=stats[1],
                amean=stats[2], ispg=ispg, next=exthd_size, cmap=b"\x00" * 4, stamp=b"\x00" * 4, rms=rms,
                labels=b"\x00" * 800)
    return MRCHeader(**zero)


def make_header(shape, cella, cellb, mz=1, dtype=np.float32, order=(1, 2, 3), dmin=0, dmax=-1, dmean=-2, rms=-1,
                exthd_size=0, ispg=0):
    """Header for a (nz, ny, nx) volume; the defaults dmax < dmin, dmean < both and rms < 0 mark the statistics as
    not computed (MRC2014 convention)."""
    return _header(shape[2], shape[1], shape[0], get_mode(dtype), mz, cella, cellb, (dmin, dmax, dmean), rms, ispg,
                   exthd_size)


def write(f, array, header=None, extended_header=b"", ax=1, ay=1, az=1, alpha=0, beta=0, gamma=0):
    """Write header + extended header + voxels.  Without a header a mode-2 (float32) one is built from the array
    (3-D, (nz, ny, nx)) with its min / max / mean / std."""
    if header is None:
        header = _header(array.shape[2], array.shape[1], array.shape[0], 2, 1, (ax, ay, az), (alpha, beta, gamma),
                         (array.min(), array.max(), array.mean()), array.std(), 0, len(extended_header))
    f.write(header_struct.pack(*header))
    f.write(extended_header)
    f.write(array.tobytes())
