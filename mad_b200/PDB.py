"""PDB with the part of the reference's interface that sits in front of the hot path
(mad/PDB.py:8-80, 131-209, 215-292): fixed-column coordinate reader and ``structure_to_density``.

The density simulation (mass-weighted trilinear splat, Gaussian of sigma = resolution / (pi sqrt 2) /
voxelsp truncated at 3 sigma, max normalisation, isovalue cut) runs in libmad_b200.so; the result can stay
on the device (``structure_to_density_device``) so that ``MapSpace`` in PDB mode starts from HBM.
The manipulation helpers of the reference's class (rotate_atoms, translate_atoms, rmsd, write_pdb; mad/PDB.py:80-128)
are host glue with the reference's semantics, used by ``structure_utils``.
"""
import ctypes as C
import os
import sys
from math import ceil, floor, sqrt

import numpy as np
import torch

from . import _lib
from . import mrc as _mrc
from ._lib import call

MASS = {"H": 1.00797, "BE": 9.01218, "C": 12.011, "N": 14.0067, "O": 15.9994, "F": 18.998403, "S": 32.06, "P": 30.97376,
        "MG": 24.305, "CL": 35.453, "K": 39.0983, "CA": 40.078, "MN": 54.9380, "FE": 55.847, "NI": 58.70, "CU": 63.546,
        "ZN": 65.38, "SE": 78.96}                                        # mad/PDB.py:220-221


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def write_situs(outname, grid, voxelsp, dxi, dyi, dzi):
    """Situs text map exactly as the reference writes it (mad/PDB.py:165-179): header line, blank line, then the voxels
    x-fastest, a line break before every 10th value."""
    dxb, dyb, dzb = grid.shape
    with open(outname, "w") as f:
        f.write("%f %f %f %f %i %i %i\n\n" % (voxelsp, dxi, dyi, dzi, dxb, dyb, dzb))
        voxi = 0
        for z in range(dzb):
            for y in range(dyb):
                for x in range(dxb):
                    if (voxi + 1) % 10 == 0:
                        f.write("\n")
                    f.write("   %6.6f   " % grid[x][y][z])
                    voxi += 1


class PDB(object):
    def __init__(self, pdb_file):
        self.pdb_file = pdb_file
        self.coords = []
        self.info = []
        if not os.path.exists(self.pdb_file):
            print("PDB> File not found: %s" % self.pdb_file)
            sys.exit(1)
        self.CA_idx = []
        self.BB_idx = []
        c = 0
        at_num = res_num = 0
        at_name = res_name = chain_id = element_symbol = ""
        x = y = z = 0.0
        with open(self.pdb_file, "r") as pdb:
            for line in pdb:
                line_type = line[0:6].strip()
                if line_type in ["ATOM", "HETATM"]:
                    try:                                                  # mad/PDB.py:46-58 (a bad line repeats the previous values)
                        at_num = int(line[6:11].strip())
                        at_name = line[12:16].strip()
                        res_name = line[17:20]
                        chain_id = line[21]
                        res_num = int(line[22:26].strip())
                        x = float(line[30:38])
                        y = float(line[38:46])
                        z = float(line[46:54])
                        element_symbol = line[76:78].strip()
                    except Exception:
                        pass
                    self.info.append([at_num, at_name, res_name, chain_id, res_num, element_symbol, line_type])
                    self.coords.append([x, y, z])
                    if at_name == "CA":
                        self.CA_idx.append(c)
                    if at_name in ["C", "CA", "N", "O"]:
                        self.BB_idx.append(c)
                    c += 1
        self.coords = np.array(self.coords)
        self.CA_idx = tuple(self.CA_idx)
        self.n_atoms = len(self.coords)
        self.n_CA = len(self.CA_idx)
        self.minx, self.maxx = np.amin(self.coords[:, 0]), np.amax(self.coords[:, 0])
        self.miny, self.maxy = np.amin(self.coords[:, 1]), np.amax(self.coords[:, 1])
        self.minz, self.maxz = np.amin(self.coords[:, 2]), np.amax(self.coords[:, 2])

    def get_coords(self):
        return self.coords

    def set_coords(self, coords):
        self.coords = coords.copy()

    def rotate_atoms(self, rot_mat):
        self.coords = np.dot(self.coords, rot_mat)                         # mad/PDB.py:109-110

    def translate_atoms(self, trans_vec):
        self.coords += np.array(trans_vec)                                 # mad/PDB.py:112-113

    def get_rmsd_with(self, pdb):
        d2 = np.square(self.coords - pdb.coords)                           # mad/PDB.py:115-117
        return np.sqrt(np.sum(d2, axis=(0, 1)) / d2.shape[0])

    def get_rmsdCA_with(self, pdb):
        if not len(self.CA_idx):                                           # mad/PDB.py:119-128
            print("PDB> No alpha carbons detected; returning all-atom RMSD instead.")
            return self.get_rmsd_with(pdb)
        d2 = np.square(self.coords[self.CA_idx, :] - pdb.coords[pdb.CA_idx, :])
        return np.sqrt(np.sum(d2, axis=(0, 1)) / d2.shape[0])

    def write_pdb(self, outname):
        """Fixed-column ATOM / HETATM records, occupancy 1.00, B 0.00 (mad/PDB.py:80-94): 4-letter atom names start
        one column earlier."""
        with open(outname, "w") as f:
            for rec, (x, y, z) in zip(self.info, self.coords):
                at_num, at_name, res_name, chain_id, res_num, element, line_type = rec
                name = " %-4s " % at_name if len(at_name) == 4 else "  %-3s " % at_name
                f.write("%-6s%5i%s%3s%2s%4s    %8.3f%8.3f%8.3f%6.2f%6.2f          %-2s\n"
                        % (line_type, at_num, name, res_name, chain_id, res_num, x, y, z, 1.0, 0.0, element))

    # ---- density ----------------------------------------------------------------------------------
    def _masses(self):
        out = np.empty(len(self.info), dtype=np.float64)
        for i, rec in enumerate(self.info):
            elem = rec[-2].upper()
            if elem in MASS:
                out[i] = MASS[elem]
            else:
                print("PDB> (dens) Element %s not in dict. Using mass of carbon." % elem)
                out[i] = MASS["C"]
        return out

    def structure_to_density_device(self, resolution, voxelsp, isovalue=0.0, pad=0):
        """(float32 CUDA tensor [x][y][z], dxi, dyi, dzi): mad/PDB.py:131-163 on the device."""
        if not torch.cuda.is_available():
            raise _lib.MadError("mad_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        st = _stream()
        xyz = np.ascontiguousarray(self.coords, dtype=np.float64)
        mass = self._masses()
        # lattice registration and box size (host scalars), mad/PDB.py:236-254
        minx = voxelsp * floor(np.amin(xyz[:, 0]) / voxelsp)
        maxx = voxelsp * ceil(np.amax(xyz[:, 0]) / voxelsp)
        miny = voxelsp * floor(np.amin(xyz[:, 1]) / voxelsp)
        maxy = voxelsp * ceil(np.amax(xyz[:, 1]) / voxelsp)
        minz = voxelsp * floor(np.amin(xyz[:, 2]) / voxelsp)
        maxz = voxelsp * ceil(np.amax(xyz[:, 2]) / voxelsp)
        margin = 2 + pad
        pxb = ceil((maxx - minx) / voxelsp) + 2 * margin + 1
        pyb = ceil((maxy - miny) / voxelsp) + 2 * margin + 1
        pzb = ceil((maxz - minz) / voxelsp) + 2 * margin + 1
        d_xyz = torch.from_numpy(xyz).to(dev)
        d_mass = torch.from_numpy(mass).to(dev)
        grid = torch.empty((pxb, pyb, pzb), dtype=torch.float64, device=dev)
        mins = np.array([minx, miny, minz], dtype=np.float64)
        call("mad_density_splat", _ptr(d_xyz), _ptr(d_mass), len(xyz), mins.ctypes.data_as(C.c_void_p), C.c_double(voxelsp),
             margin, pxb, pyb, pzb, _ptr(grid), st)
        scratch = torch.empty(1, dtype=torch.int64, device=dev)
        call("mad_normalise_f64", _ptr(grid), grid.numel(), _ptr(scratch), st)
        # Gaussian, mad/PDB.py:138-150: the 3-D kernel exp(-(x^2+y^2+z^2) / 2 sig^2) / sum is a product of 1-D kernels
        sig = resolution / (np.pi * sqrt(2)) / voxelsp
        r = int(ceil(3.0 * sig))
        t = np.arange(-r, r + 1)
        g1 = np.exp(-(t * t) / (2.0 * sig ** 2))
        w = torch.from_numpy(np.ascontiguousarray(g1 / g1.sum())).to(dev)
        a = torch.empty((pxb + 2 * r, pyb, pzb), dtype=torch.float64, device=dev)
        call("mad_conv_full_f64", _ptr(grid), 1, pxb, pyb * pzb, _ptr(w), r, _ptr(a), 0, st)
        b = torch.empty((pxb + 2 * r, pyb + 2 * r, pzb), dtype=torch.float64, device=dev)
        call("mad_conv_full_f64", _ptr(a), pxb + 2 * r, pyb, pzb, _ptr(w), r, _ptr(b), 0, st)
        dens = torch.empty((pxb + 2 * r, pyb + 2 * r, pzb + 2 * r), dtype=torch.float32, device=dev)
        call("mad_conv_full_f64", _ptr(b), (pxb + 2 * r) * (pyb + 2 * r), pzb, 1, _ptr(w), r, _ptr(dens), 1, st)
        dxi = minx - (r + margin) * voxelsp
        dyi = miny - (r + margin) * voxelsp
        dzi = minz - (r + margin) * voxelsp
        # [0, 1] range, isovalue cut (float32), mad/PDB.py:158-159
        mx = torch.zeros(1, dtype=torch.int32, device=dev)
        call("mad_grid_max", _ptr(dens), dens.numel(), _ptr(mx), st)
        vmax = float(_lib.lib.mad_grid_max_decode(C.c_uint(int(mx.item()) & 0xFFFFFFFF)))
        call("mad_threshold_normalise", _ptr(dens), dens.numel(), C.c_float(-np.inf), C.c_float(vmax), 1, st)
        if isovalue:
            call("mad_threshold_normalise", _ptr(dens), dens.numel(), C.c_float(isovalue), C.c_float(1.0), 0, st)
        return dens, dxi, dyi, dzi

    def structure_to_density(self, resolution, voxelsp, isovalue=0.0, pad=0, outname=""):
        """Reference signature and return value: (grid float32 [x][y][z], dxi, dyi, dzi)."""
        dens, dxi, dyi, dzi = self.structure_to_density_device(resolution, voxelsp, isovalue, pad)
        grid = dens.cpu().numpy()
        if outname != "":
            ext = os.path.splitext(outname)[-1].lower()
            if ext in [".sit", ".situs"]:
                write_situs(outname, grid, voxelsp, dxi, dyi, dzi)
            else:
                _mrc.write_mrc(outname, grid.transpose(2, 1, 0), voxelsp, origin=(dxi, dyi, dzi))
        return grid.astype(np.float32), dxi, dyi, dzi
