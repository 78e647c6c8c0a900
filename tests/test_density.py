"""Atoms -> density (mad/PDB.py:131-163, 215-292): oracle restatement against the
reference-generated fixture (CPU) and the device implementation against both (GPU).

Tolerance.  The reference convolves with scipy.signal.convolve (FFT or direct sum chosen by size) and splats
the atoms sequentially in float64; the device splats with float64 atomics and convolves as three 1-D passes.
Both are float64 computations of the same quantity, so the float32 maps agree to one float32 ulp of the
maximum (1.2e-7); isovalue-cut voxels may flip only where the density is within that distance of the cut."""
import os

import numpy as np
import pytest

import helpers as H

CASES = (("a", dict(resolution=8.0, voxelsp=2.0)), ("b", dict(resolution=4.0, voxelsp=1.0, isovalue=0.2)),
         ("c", dict(resolution=5.0, voxelsp=2.0, isovalue=0.2, pad=1)))


def _atoms(g):
    coords, masses = [], []
    mass = {"C": 12.011, "N": 14.0067, "O": 15.9994}
    for line in bytes(g["pdb_text"]).decode().splitlines():
        if line[:6].strip() in ("ATOM", "HETATM"):
            coords.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            masses.append(mass.get(line[76:78].strip().upper(), 12.011))
    return np.array(coords), np.array(masses)


@pytest.mark.parametrize("tag,kw", CASES)
def test_oracle_density_equals_reference(tag, kw):
    import density_oracle as do
    g = H.golden("density")
    coords, masses = _atoms(g)
    grid, dxi, dyi, dzi = do.structure_to_density(coords, masses, **kw)
    assert grid.shape == g[tag + "_grid"].shape and grid.dtype == np.float32
    assert np.array_equal([dxi, dyi, dzi], g[tag + "_origin"])
    assert np.abs(grid - g[tag + "_grid"]).max() <= 1.2e-7      # scipy.signal.convolve may pick FFT or direct sums


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kw", CASES)
def test_device_density_equals_reference(tmp_path, tag, kw):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mad_b200.PDB import PDB
    g = H.golden("density")
    path = os.path.join(str(tmp_path), "case.pdb")
    open(path, "wb").write(bytes(g["pdb_text"]))
    pdb = PDB(path)
    assert pdb.n_atoms == 300
    grid, dxi, dyi, dzi = pdb.structure_to_density(**kw)
    ref = g[tag + "_grid"]
    assert grid.shape == ref.shape and grid.dtype == np.float32
    assert np.array_equal([dxi, dyi, dzi], g[tag + "_origin"])
    iso = kw.get("isovalue", 0.0)
    near_cut = np.abs(np.where(ref == 0, grid, ref) - iso) <= 2.4e-7 if iso else np.zeros(ref.shape, bool)
    assert np.abs(grid - ref)[~near_cut].max() <= 1.2e-7
    assert near_cut.sum() <= 2
    assert grid.max() == 1.0
    dev, *_ = pdb.structure_to_density_device(**kw)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), grid)


def test_situs_writer_equals_reference_and_reads_back(tmp_path):
    """``structure_to_density(outname="*.sit")``: the written text equals the reference's byte for byte
    (mad/PDB.py:165-179) and this package's own readers (MapSpace / Dmap header + voxel parse) accept it."""
    from mad_b200.PDB import write_situs
    from mad_b200.MapSpace import MapSpace
    g = H.golden("density")
    grid, org = g["c_grid"], g["c_origin"]
    path = os.path.join(str(tmp_path), "out.sit")
    write_situs(path, grid, 2.0, float(org[0]), float(org[1]), float(org[2]))
    assert open(path, "rb").read() == bytes(g["c_sit_text"])
    ms = MapSpace(path)
    back, origin = ms._load()                                     # host-side reader only (no device work)
    assert back.shape == grid.shape and tuple(origin) == tuple(float(o) for o in org) and ms.voxelsp == 2.0
    want = np.round(grid.astype(np.float64), 6)                   # "%6.6f" text, then max-normalised by the reader
    assert np.abs(back - (want / want.max()).astype(np.float32)).max() <= 1e-6
