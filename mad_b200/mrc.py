"""Minimal MRC2014 reader / writer (mode 2, float32) for the map formats MaD reads and writes
(mad/MapSpace.py:97-114, mad/Dmap.py:26-43,392-415, mad/PDB.py:181-206).  Replaces the `mrcfile`
dependency with the subset of header fields the reference touches."""
import struct

import numpy as np

_MODE_DTYPE = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16, 12: np.float16}


class MrcHeader(object):
    pass


def read_mrc(path):
    """Returns (header, data[z][y][x]-style array of shape (ns, nr, nc))."""
    with open(path, "rb") as f:
        raw = f.read(1024)
        if len(raw) < 1024:
            raise ValueError("MRC header truncated: %s" % path)
        stamp = raw[212:216]
        end = ">" if stamp[:1] in (b"\x11",) else "<"
        i = struct.unpack(end + "10i", raw[0:40])
        h = MrcHeader()
        h.nx, h.ny, h.nz, h.mode, h.nxstart, h.nystart, h.nzstart, h.mx, h.my, h.mz = i
        h.cella = struct.unpack(end + "3f", raw[40:52])
        h.mapc, h.mapr, h.maps = struct.unpack(end + "3i", raw[64:76])
        h.nsymbt = struct.unpack(end + "i", raw[92:96])[0]
        h.origin = struct.unpack(end + "3f", raw[196:208])
        if h.mode not in _MODE_DTYPE:
            raise ValueError("unsupported MRC mode %d in %s" % (h.mode, path))
        f.seek(1024 + max(h.nsymbt, 0))
        dt = np.dtype(_MODE_DTYPE[h.mode]).newbyteorder(end)
        data = np.fromfile(f, dtype=dt, count=h.nx * h.ny * h.nz).reshape(h.nz, h.ny, h.nx)
    h.voxel_size = tuple(np.float32(c / m) if m else np.float32(0) for c, m in zip(h.cella, (h.mx, h.my, h.mz)))
    return h, data


def write_mrc(path, data_szyx, voxelsp, origin=(0.0, 0.0, 0.0), nstart=(0, 0, 0)):
    """data_szyx: float32 array indexed [section][row][column] (= grid.transpose(2, 1, 0))."""
    data = np.ascontiguousarray(data_szyx, dtype="<f4")
    nz, ny, nx = data.shape
    hdr = bytearray(1024)
    struct.pack_into("<10i", hdr, 0, nx, ny, nz, 2, nstart[0], nstart[1], nstart[2], nx, ny, nz)
    struct.pack_into("<6f", hdr, 40, nx * voxelsp, ny * voxelsp, nz * voxelsp, 90.0, 90.0, 90.0)
    struct.pack_into("<3i", hdr, 64, 1, 2, 3)
    struct.pack_into("<3f", hdr, 76, float(data.min()), float(data.max()), float(data.mean()))
    struct.pack_into("<3f", hdr, 196, float(origin[0]), float(origin[1]), float(origin[2]))
    hdr[208:212] = b"MAP "
    hdr[212:216] = b"\x44\x44\x00\x00"
    struct.pack_into("<f", hdr, 216, float(data.std()))
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(data.tobytes())
