"""structure_utils with the reference's interface (mad/structure_utils.py:8-259): ``refine_pdb`` and ``get_overlap``
on the device, ``move_structure`` / ``move_copy_structure`` as host glue.

``refine_pdb`` runs the whole rigid-body steepest-ascent loop in ONE kernel launch (``refine_kernel``, score.cu), one
CTA per pose; ``refine_poses`` is the batched entry point for the candidate loop of ``MaD._refine_filtered_solutions``
(mad/MaD.py:556-577), which the reference refines one at a time.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import call
from .Dmap import Dmap
from .PDB import PDB
from .math_utils import euler_rod_mat


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def move_structure(original_struct, t=None, a=0.375, b=1.735, c=2.452, suffix=""):
    """mad/structure_utils.py:8-28."""
    moved_struct = original_struct.replace(".pdb", "_moved%s.pdb" % suffix)
    pdb = PDB(original_struct)
    for axis, ang in (([1, 0, 0], a), ([0, 1, 0], b), ([0, 0, 1], c)):
        pdb.rotate_atoms(euler_rod_mat(axis, ang))
    if t is None:
        pdb.translate_atoms(-np.mean(pdb.get_coords(), axis=0))
    else:
        pdb.translate_atoms(t)
    pdb.write_pdb(moved_struct)
    return moved_struct


def move_copy_structure(original_struct, moved_struct, transform=False, t=[150, 0, 0], a=0.375, b=1.735, c=2.452):
    """mad/structure_utils.py:30-56."""
    pdb = PDB(original_struct)
    if transform:
        for axis, ang in (([1, 0, 0], a), ([0, 1, 0], b), ([0, 0, 1], c)):
            pdb.rotate_atoms(euler_rod_mat(axis, ang))
        pdb.translate_atoms(-np.mean(pdb.get_coords(), axis=0))
        if len(t):
            pdb.translate_atoms(t)
    pdb.write_pdb(moved_struct)
    return moved_struct


class RefineField(object):
    """Gradient field of a map prepared once for many refinements: np.gradient as float4 per voxel in HBM and the
    axis coordinates of mad/structure_utils.py:75-77."""

    def __init__(self, dmap):
        if not torch.cuda.is_available():
            raise _lib.MadError("mad_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if isinstance(dmap, Dmap):
            grid = dmap.device_grid()
        else:
            grid = torch.from_numpy(np.ascontiguousarray(dmap.grid3d, dtype=np.float32)).cuda()
        self.voxsp = float(dmap.voxsp)
        self.shape = [int(v) for v in grid.shape]
        sx, sy, sz = self.shape
        self.grad = torch.empty((sx, sy, sz, 4), dtype=torch.float32, device=grid.device)
        call("mad_gradient", _ptr(grid), sx, sy, sz, _ptr(self.grad), _stream())
        v = dmap.voxsp
        pts = [np.arange(o, v * n + o, v)[:b] for o, n, b in ((dmap.xi, sx, dmap.xb), (dmap.yi, sy, dmap.yb), (dmap.zi, sz, dmap.zb))]
        if [len(p) for p in pts] != self.shape:
            raise ValueError("There are %d points and %d values in dimension 0" % (len(pts[0]), sx))   # as RGI raises
        self.points = [torch.from_numpy(np.ascontiguousarray(p, dtype=np.float64)).to(grid.device) for p in pts]


def refine_poses(field, poses, n_steps=500, max_step_size=0.5, min_step_size=0.01):
    """Refines P poses [P][n][3] (float64) of one structure in one launch.
    Returns (coords [P][n][3] float64 NumPy, converged bool [P], step int [P], nan bool [P])."""
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    if poses.ndim == 2:
        poses = poses[None]
    n_p, n_atoms, _ = poses.shape
    dev = field.grad.device
    center = np.mean(poses, axis=1)                                        # pdb_center, :66
    max_dist = np.amax(np.linalg.norm(poses - center[:, None, :], axis=2), axis=1)   # :67
    d_init = torch.from_numpy(poses).to(dev)
    d_center = torch.from_numpy(np.ascontiguousarray(center)).to(dev)
    d_max = torch.from_numpy(np.ascontiguousarray(max_dist)).to(dev)
    d_out = torch.empty_like(d_init)
    d_meta = torch.empty((n_p, 4), dtype=torch.float64, device=dev)
    sx, sy, sz = field.shape
    call("mad_refine_rigid", _ptr(field.grad), sx, sy, sz, _ptr(field.points[0]), _ptr(field.points[1]), _ptr(field.points[2]),
         C.c_double(field.voxsp), _ptr(d_init), _ptr(d_center), _ptr(d_max), n_p, n_atoms, int(n_steps),
         C.c_double(max_step_size), C.c_double(min_step_size), _ptr(d_out), _ptr(d_meta), _stream())
    meta = d_meta.cpu().numpy()
    return d_out.cpu().numpy(), meta[:, 0] != 0, meta[:, 1].astype(np.int64), meta[:, 2] != 0


def refine_pdb(dmap, pdb, n_steps=500, max_step_size=0.5, min_step_size=0.01, idx=-1):
    """mad/structure_utils.py:58-161: moves ``pdb.coords`` in place; returns (rmsd_beforeAfter, converged, step)."""
    field = dmap if isinstance(dmap, RefineField) else RefineField(dmap)
    init = np.array(pdb.coords, dtype=np.float64)
    coords, conv, step, bad = refine_poses(field, init, n_steps, max_step_size, min_step_size)
    pdb.set_coords(coords[0])
    if bad[0]:
        return np.nan, False, int(step[0])
    if len(pdb.CA_idx):                                                    # :154-158
        d2 = np.square(pdb.coords[pdb.CA_idx, :] - init[pdb.CA_idx, :])
    else:
        d2 = np.square(pdb.coords - init)
    return np.sqrt(np.sum(d2, axis=(0, 1)) / d2.shape[0]), bool(conv[0]), int(step[0])


def get_overlap(g1, g2, voxsp, isovalue=1e-8):
    """mad/structure_utils.py:163-259: g = (grid, xi, yi, zi); share of g1's voxels > 0 that are > 0 in g2 over the
    common box.  Both grids are cut at the isovalue first (CUDA tensors in place, NumPy inputs on their device copy)."""
    grid1, xi1, yi1, zi1 = g1
    grid2, xi2, yi2, zi2 = g2
    if not torch.cuda.is_available():
        raise _lib.MadError("mad_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    d1 = Dmap._as_device_grid(grid1, dev)
    d2 = Dmap._as_device_grid(grid2, dev)
    for d in (d1, d2):
        call("mad_threshold_normalise", _ptr(d), d.numel(), C.c_float(isovalue), C.c_float(1.0), 0, _stream())
    m1 = Dmap.__new__(Dmap)
    m1._dev, m1._host, m1.voxsp = d1, None, voxsp
    m1.xi, m1.yi, m1.zi = xi1, yi1, zi1
    m1.xb, m1.yb, m1.zb = [int(v) for v in d1.shape]
    s = m1._box_scores(d2, xi2, yi2, zi2, 0.0)
    if s is None:
        return 0
    m1_vals = m1._count_gt(d1, 0.0)
    if m1_vals == 0:
        return 0
    return int(s[6]) / m1_vals
