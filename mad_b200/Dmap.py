"""Dmap with the reference's interface (mad/Dmap.py:6-97), the voxel grid resident in HBM.

    dm = Dmap(map_name, isovalue=0.0, normalize=True, pad=0)
    dm.reduce_void(); dm.pad_grid(5); ms = MapSpace.from_grid(dm.device_grid(), dm.voxsp, (dm.xi, dm.yi, dm.zi))

File parsing is host glue as in the reference (Situs text, MRC through mad_b200.mrc); the isovalue
cut, the max normalisation, the bounding box of ``reduce_void`` and crop / zero padding run in
libmad_b200.so on the device grid (row a0 of SURVEY.md section 8).  ``grid3d`` is a lazily synchronised
NumPy copy for Python consumers; ``device_grid()`` hands the CUDA tensor to ``MapSpace.from_grid``
without a host round trip.  The scoring / masking methods of the reference's Dmap
(``mad/Dmap.py:99-377``) are outside the hot path and not provided.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

from . import _lib
from . import mrc as _mrc
from ._lib import call


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Dmap(object):
    def __init__(self, map_name, isovalue=0.0, normalize=True, pad=0):
        if not os.path.isfile(map_name):
            print("Dmap> ERROR: file %s not found" % map_name)
            sys.exit(1)
        ext = os.path.splitext(map_name)[-1]
        if ext.lower() in [".sit", ".situs"]:
            with open(map_name, "r") as f:
                header = f.readline().replace("\n", "").replace("  ", "").split(" ")
                f.readline()
                self.voxsp, self.xi, self.yi, self.zi = [float(x) for x in header[:4]]
                self.xb, self.yb, self.zb = [int(x) for x in header[4:]]
                grid1d = np.array(f.read().split(), dtype=np.float64).astype(np.float32)   # as np.fromstring(...).astype
                grid = np.reshape(grid1d, (self.xb, self.yb, self.zb), order="F")
        elif ext.lower() in [".map", ".mrc"]:
            h, data = _mrc.read_mrc(map_name)
            axis_order = [h.mapc - 1, h.mapr - 1, h.maps - 1]
            self.axis_order = axis_order
            self.voxsp = h.voxel_size[0]
            if np.all([h.nxstart, h.nystart, h.nzstart]):                  # mad/Dmap.py:34-36
                origin = np.array([h.nxstart, h.nystart, h.nzstart], dtype=int)
                self.xi, self.yi, self.zi = [origin[a] * self.voxsp for a in axis_order]
            else:                                                           # int-truncated origin, :38-39
                origin = np.array(h.origin, dtype=int)
                self.xi, self.yi, self.zi = [origin[a] for a in axis_order]
            boxdim = np.array([h.mx, h.my, h.mz], dtype=int)
            self.xb, self.yb, self.zb = [int(boxdim[a]) for a in axis_order]
            grid = np.transpose(data.copy(), axis_order[::-1])
        else:
            print("Dmap> ERROR: incompatible extension for map %s" % map_name)
            return
        self.map_name = map_name
        self.name = map_name.split("/")[-1].split(".")[0]
        self._init_from_array(grid, isovalue, normalize, pad)

    @classmethod
    def from_array(cls, grid, voxsp, origin=(0.0, 0.0, 0.0), isovalue=0.0, normalize=True, pad=0, name="grid"):
        """Array entry point (NumPy or CUDA tensor, float32 [x][y][z]); same post-processing as a file."""
        self = cls.__new__(cls)
        self.voxsp = voxsp
        self.xi, self.yi, self.zi = origin
        self.map_name = name
        self.name = name
        self.xb, self.yb, self.zb = [int(v) for v in grid.shape]
        self._init_from_array(grid, isovalue, normalize, pad)
        return self

    # ---- device side ------------------------------------------------------------------------------
    def _init_from_array(self, grid, isovalue, normalize, pad):
        if not torch.cuda.is_available():
            raise _lib.MadError("mad_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if isinstance(grid, np.ndarray):
            grid = torch.from_numpy(np.ascontiguousarray(grid, dtype=np.float32)).cuda()
        self._dev = grid.contiguous().float().clone()
        self._host = None
        vmax = self._max()
        if vmax > isovalue:                                                 # mad/Dmap.py:49-54
            thr = float(isovalue)
        else:
            print("Dmap> WARNING: asked isovalue is larger than maximum density found in file (%f). Considering isovalue=0" % vmax)
            thr = 0.0
            vmax = max(vmax, 0.0)
        call("mad_threshold_normalise", _ptr(self._dev), self._dev.numel(), C.c_float(thr), C.c_float(1.0), 0, _stream())
        if pad:
            self.pad_grid(pad)
        if np.isclose(vmax, 0):
            print("Dmap> WARNING: Max value in map is 0")
        if normalize:                                                       # mad/Dmap.py:66-67
            call("mad_threshold_normalise", _ptr(self._dev), self._dev.numel(), C.c_float(-np.inf), C.c_float(vmax), 1, _stream())

    def _max(self):
        out = torch.zeros(1, dtype=torch.int32, device=self._dev.device)
        call("mad_grid_max", _ptr(self._dev), self._dev.numel(), _ptr(out), _stream())
        return float(_lib.lib.mad_grid_max_decode(C.c_uint(int(out.item()) & 0xFFFFFFFF)))

    def device_grid(self):
        """The float32 [x][y][z] CUDA tensor (no copy)."""
        return self._dev

    @property
    def grid3d(self):
        if self._host is None:
            self._host = self._dev.cpu().numpy()
        return self._host

    @grid3d.setter
    def grid3d(self, value):
        self._dev = torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32)).cuda()
        self._host = None
        self.xb, self.yb, self.zb = [int(v) for v in self._dev.shape]

    def _crop_pad(self, x0, y0, z0, cx, cy, cz, pad):
        nx, ny, nz = [int(v) for v in self._dev.shape]
        out = torch.empty((cx + 2 * pad, cy + 2 * pad, cz + 2 * pad), dtype=torch.float32, device=self._dev.device)
        call("mad_crop_pad3d", _ptr(self._dev), nx, ny, nz, x0, y0, z0, cx, cy, cz, int(pad), _ptr(out), _stream())
        self._dev = out
        self._host = None
        self.xb, self.yb, self.zb = [int(v) for v in out.shape]

    def reduce_void(self, zeros_padding=10):
        """Crop to the bounding box of the non-zero voxels, then zero-pad (mad/Dmap.py:73-90)."""
        nx, ny, nz = [int(v) for v in self._dev.shape]
        bbox = torch.empty(6, dtype=torch.int32, device=self._dev.device)
        call("mad_grid_bbox", _ptr(self._dev), nx, ny, nz, _ptr(bbox), _stream())
        b = [int(v) for v in bbox.tolist()]
        if b[3] < 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")   # as np.amin does
        minx, miny, minz, maxx, maxy, maxz = b
        self.xi = self.xi + minx * self.voxsp
        self.yi = self.yi + miny * self.voxsp
        self.zi = self.zi + minz * self.voxsp
        self._crop_pad(minx, miny, minz, maxx - minx + 1, maxy - miny + 1, maxz - minz + 1, 0)
        self.pad_grid(zeros_padding)

    def pad_grid(self, pad):
        """np.pad(grid3d, pad) with the origin moved (mad/Dmap.py:92-97)."""
        nx, ny, nz = [int(v) for v in self._dev.shape]
        self._crop_pad(0, 0, 0, nx, ny, nz, int(pad))
        self.xi -= pad * self.voxsp
        self.yi -= pad * self.voxsp
        self.zi -= pad * self.voxsp

    def write_to_mrc(self, outname):
        """mad/Dmap.py:392-415: data as [z][y][x], nstart 0, origin = (xi, yi, zi), cella = box * voxsp."""
        _mrc.write_mrc(outname, self.grid3d.transpose(2, 1, 0), float(self.voxsp), origin=(self.xi, self.yi, self.zi))
