"""MapSpace with the reference's interface (mad/MapSpace.py:12-214), computed on the B200.

    ms = MapSpace(structure_file, resolution=0, voxelsp=0, isovalue=0.0, map_padding=9,
                  oct_mode="both", sig_init=2, sig_presmooth=1)
    ms.build_space()

File parsing and the isovalue cut are host glue (as in the reference); everything from the
zero padding on (upsampling + presmoothing, LoG, Gaussian, gradient: a1-a4) runs in
libmad_b200.so and stays resident in HBM (``ms.space``).  The reference's attributes
``map_space, grid_list, gauss_list, grad_list, rgi_space`` are served as lazily copied NumPy
views for Python consumers; Detector / Orientator / Descriptor of this package read the device
arrays directly.  ``MapSpace.from_grid`` starts from an array (NumPy or torch CUDA tensor, e.g.
``Dmap.device_grid()``) instead of a file.
"""
import os
import sys

import numpy as np

from . import mrc as _mrc
from . import pipeline as _P


class MapSpace(object):
    def __init__(self, structure_file, resolution=0, voxelsp=0, isovalue=0.0, map_padding=9, oct_mode="both",
                 sig_init=2, sig_presmooth=1):
        self.structure_file = structure_file
        self.isovalue = isovalue
        self.map_padding = map_padding
        self.PDB_mode = False
        self.name = os.path.splitext(os.path.split(structure_file)[-1])[0]
        self.sig_init = sig_init
        self.sig_presmooth = sig_presmooth
        self.oct_mode = oct_mode
        self.exact_f64 = True
        self._grid = None
        self._cache = {}
        self.space = None
        if oct_mode not in ["base", "up", "both"]:
            print("MaD> WARNING: #octave not set properly (%s), reverting to 'base'" % oct_mode)
            self.oct_mode = "base"
        self.ext = os.path.splitext(structure_file)[-1].lower()
        if self.ext == ".pdb":
            self.PDB_mode = True
            self.voxelsp = voxelsp
            if self.voxelsp == 0:
                print("MaD> ERROR: if providing a PDB, voxel spacing is mandatory")
                sys.exit(1)
            self.resolution = resolution
            if self.resolution == 0:
                print("MaD> ERROR: if providing a PDB, resolution is mandatory")
                sys.exit(1)
        elif self.ext not in [".situs", ".sit", ".map", ".mrc", ".grid"]:
            print("MaD> ERROR: please provide a valid structure file (pdb, sit, situs, map or mrc format)")
            print(self.ext)
            sys.exit(1)

    @classmethod
    def from_grid(cls, grid, voxelsp, origin=(0.0, 0.0, 0.0), name="grid", **kw):
        """Array entry point: ``grid`` float32 [x][y][z] (NumPy or CUDA tensor), already thresholded
        and normalised the way the file readers deliver it."""
        ms = cls(name + ".grid", **kw)
        ms.name = name
        ms.voxelsp = voxelsp
        ms._grid = grid
        ms._origin = tuple(float(o) for o in origin)
        return ms

    # ---- host glue: the three readers of mad/MapSpace.py:73-114 -------------------------------
    def _load(self):
        if self._grid is not None:
            return self._grid, self._origin
        if self.PDB_mode:                                     # mad/MapSpace.py:73-76: simulated density of the structure
            from .PDB import PDB
            grid, xi, yi, zi = PDB(self.structure_file).structure_to_density_device(self.resolution, self.voxelsp,
                                                                                    isovalue=self.isovalue)
            return grid, (xi, yi, zi)
        if self.ext in [".situs", ".sit"]:
            with open(self.structure_file, "r") as sit:
                header = sit.readline().replace("\n", "").replace("  ", "").split(" ")
                sit.readline()
                grid1d = np.array(sit.read().split(), dtype=np.float64)
            self.voxelsp, xi, yi, zi = [float(x) for x in header[:4]]
            xb, yb, zb = [int(x) for x in header[4:]]
            grid1d[grid1d < self.isovalue] = 0
            grid = np.reshape(grid1d, (xb, yb, zb), order="F")
            grid = grid / np.amax(grid).astype(np.float32)
            return grid.astype(np.float32), (xi, yi, zi)
        h, data = _mrc.read_mrc(self.structure_file)
        axis_order = [h.mapc - 1, h.mapr - 1, h.maps - 1]
        self.voxelsp = h.voxel_size[0]
        if all([h.nxstart, h.nystart, h.nzstart]):
            origin = np.array([h.nxstart, h.nystart, h.nzstart], dtype=int)
            xi, yi, zi = [origin[a] * self.voxelsp for a in axis_order]      # reference bug D.1 fixed
        else:
            origin = np.array(h.origin).astype(int)                           # truncation kept (D.2)
            xi, yi, zi = [origin[a] for a in axis_order]
        grid = np.transpose(data.copy(), axis_order[::-1]).astype(np.float32)
        grid[grid < self.isovalue] = 0
        return grid, (xi, yi, zi)

    def build_space(self):
        print("MaD> Building map space for %s..." % self.name)
        grid, (xi, yi, zi) = self._load()
        if self.map_padding:
            xi -= self.map_padding * self.voxelsp
            yi -= self.map_padding * self.voxelsp
            zi -= self.map_padding * self.voxelsp
        self.xi, self.yi, self.zi = xi, yi, zi
        self.voxelsp_list = {"both": [self.voxelsp / 2, self.voxelsp], "up": [self.voxelsp / 2],
                             "base": [self.voxelsp]}[self.oct_mode]                     # mad/MapSpace.py:149-163
        self.space = _P.build_space(grid, self.map_padding, self.sig_init, self.sig_presmooth,
                                    exact_f64=self.exact_f64, keep_gauss=True, full_gradient=False, oct_mode=self.oct_mode)
        self._cache = {}

    # ---- NumPy views of the device arrays (lazy) ------------------------------------------------
    def _host(self, key, tensors, post=None):
        if self.space is None:
            raise AttributeError("%s: call build_space() first" % key)
        if key not in self._cache:
            out = [t.cpu().numpy() for t in tensors]
            self._cache[key] = [post(a) for a in out] if post else out
        return self._cache[key]

    @property
    def grid_list(self):
        return self._host("grid_list", self.space.grids)

    @property
    def map_space(self):
        return self._host("map_space", self.space.logs)

    @property
    def gauss_list(self):
        return self._host("gauss_list", self.space.gauss)

    @property
    def grad_list(self):
        if "grad_list" not in self._cache:
            _P.full_gradient(self.space)                        # tiles the stages have not asked for yet
        return self._host("grad_list", self.space.grad4[:self.space.n_oct], lambda a: a[..., :3])

    @property
    def rgi_space(self):
        if "rgi_space" not in self._cache:
            from scipy.interpolate import RegularGridInterpolator as RGI
            self._cache["rgi_space"] = [RGI(points=[np.arange(s) for s in g.shape[:3]], values=g, method="nearest")
                                        for g in self.grad_list]
        return self._cache["rgi_space"]
