"""Scoring reductions and rigid refinement (mad/Dmap.py:99-372, mad/structure_utils.py:58-259):
oracle restatement against the reference-generated fixture (CPU) and the device implementation against both (GPU).

Tolerances.  ``get_overlap`` and ``mask_with`` are integer / copy work: exact.  The two CCC scores are float32 BLAS
dots in the reference (summation order = whatever sdot does) and float64 fixed-order sums here: 2e-6 relative.
``refine_pdb`` is a float64 trajectory whose force / torque sums are sequential ``np.sum`` in the reference and a tree
reduction on the device; the oracle measures the sensitivity to that order (~1e-13 A), the test allows 1e-9 A on the
final coordinates and requires the same convergence flag and step count."""
import os

import numpy as np
import pytest

import helpers as H

REFINE_CASES = (("mad", dict(n_steps=500, max_step_size=1, min_step_size=0.1)), ("default", dict()),
                ("short", dict(n_steps=7, max_step_size=1, min_step_size=0.1)))


def _case():
    g = H.golden("score")
    return g, g["map_grid"], g["map_origin"], float(g["voxsp"]), g["sub_grid"], g["sub_origin"], g["far_origin"]


# ---- CPU: the oracle is pinned on the reference's own output ---------------------------------------------------
@pytest.mark.parametrize("tag,kw", REFINE_CASES)
def test_oracle_refine_equals_reference(tag, kw):
    import score_oracle as so
    g, G, O, v, *_ = _case()
    cur, rmsd, conv, step = so.refine_pdb(G.copy(), O, v, g["moved"], ca_idx=range(len(g["moved"])), **kw)
    meta = g["refine_%s_meta" % tag]
    assert np.array_equal(cur, g["refine_%s_coords" % tag])               # bit for bit
    assert rmsd == meta[0] and bool(conv) == bool(meta[1]) and step == int(meta[2])


def test_oracle_refine_sensitivity_to_summation_order():
    import score_oracle as so
    g, G, O, v, *_ = _case()
    pairwise = lambda a: np.array([np.sum(np.ascontiguousarray(a[:, i])) for i in range(3)])
    a, *_ = so.refine_pdb(G.copy(), O, v, g["moved"], n_steps=500, max_step_size=1, min_step_size=0.1)
    b, *_ = so.refine_pdb(G.copy(), O, v, g["moved"], n_steps=500, max_step_size=1, min_step_size=0.1, sum_rows=pairwise)
    assert np.abs(a - b).max() < 1e-11


def test_oracle_scores_equal_reference():
    import score_oracle as so
    g, G, O, v, S, SO, F = _case()
    for tag, iso in (("iso0", 0), ("iso2", 0.2)):
        assert abs(so.ccc_with_grid(G.copy(), O, v, S.copy(), SO, iso) - g["ccc_grid_" + tag]) < 2e-6 * g["ccc_grid_" + tag]
        assert abs(so.ccc_with_dmap(G, O, v, S, SO, iso) - g["ccc_dmap_" + tag]) < 2e-6 * g["ccc_dmap_" + tag]
    assert abs(so.ccc_with_grid(G.copy(), O, v, S.copy(), F) - g["ccc_grid_far"]) < 2e-6 * g["ccc_grid_far"]
    assert abs(so.ccc_with_dmap(G, O, v, S, F) - g["ccc_dmap_far"]) < 2e-6 * g["ccc_dmap_far"]
    assert so.overlap(G.copy(), O, S.copy(), SO, 2) == g["overlap"]
    assert so.overlap(G.copy(), O, S.copy(), F, 2) == g["overlap_far"]
    assert so.overlap(G.copy(), O, S.copy(), SO, 2, 0.3) == g["overlap_iso"]
    assert np.array_equal(so.mask_with(G.copy(), O, S, SO, v), g["masked"])
    assert np.array_equal(so.mask_with(G.copy(), O, S, F, v), g["masked_far"])
    # disjoint boxes
    assert so.ccc_with_grid(G.copy(), O, v, S.copy(), SO + 500.0) == 0
    assert so.overlap(G.copy(), O, S.copy(), SO - 500.0, 2) == 0


# ---- GPU: the device path against the reference fixture and the oracle ---------------------------------------------
def _cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.mark.gpu
def test_device_scores_equal_reference():
    _cuda()
    from mad_b200.Dmap import Dmap
    from mad_b200.structure_utils import get_overlap
    g, G, O, v, S, SO, F = _case()

    def dm(grid, org):
        return Dmap.from_array(grid.copy(), v, origin=tuple(float(x) for x in org), normalize=False)

    for tag, iso in (("iso0", 0), ("iso2", 0.2)):
        d = dm(G, O)
        got = d.get_CCC_with_grid(S.copy(), *SO, isovalue=iso)
        assert abs(got - g["ccc_grid_" + tag]) < 2e-6 * g["ccc_grid_" + tag]
        if iso:                                                             # the in-place cut of the reference
            assert np.array_equal(d.grid3d, np.where(G < iso, 0, G).astype(np.float32))
        got = dm(G, O).get_CCC_with_dmap(dm(S, SO), isovalue=iso)
        assert abs(got - g["ccc_dmap_" + tag]) < 2e-6 * g["ccc_dmap_" + tag]
    assert abs(dm(G, O).get_CCC_with_grid(S.copy(), *F) - g["ccc_grid_far"]) < 2e-6 * g["ccc_grid_far"]
    assert abs(dm(G, O).get_CCC_with_dmap(dm(S, F)) - g["ccc_dmap_far"]) < 2e-6 * g["ccc_dmap_far"]
    assert get_overlap([G.copy(), *O], [S.copy(), *SO], 2) == g["overlap"]
    assert get_overlap([G.copy(), *O], [S.copy(), *F], 2) == g["overlap_far"]
    assert get_overlap([G.copy(), *O], [S.copy(), *SO], 2, isovalue=0.3) == g["overlap_iso"]
    d = dm(G, O)
    d.mask_with(dm(S, SO))
    assert np.array_equal(d.grid3d, g["masked"])
    d = dm(G, O)
    d.mask_with(dm(S, F))
    assert np.array_equal(d.grid3d, g["masked_far"])
    # disjoint boxes and self-scores
    assert dm(G, O).get_CCC_with_grid(S.copy(), *(SO + 500.0)) == 0
    assert get_overlap([G.copy(), *O], [S.copy(), *(SO - 500.0)], 2) == 0
    assert abs(dm(G, O).get_CCC_with_grid(G.copy(), *O) - 1.0) < 1e-15
    assert get_overlap([G.copy(), *O], [G.copy(), *O], 2) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kw", REFINE_CASES)
def test_device_refine_equals_reference(tmp_path, tag, kw):
    _cuda()
    from mad_b200.Dmap import Dmap
    from mad_b200.PDB import PDB
    from mad_b200.structure_utils import refine_pdb
    g, G, O, v, *_ = _case()
    path = os.path.join(str(tmp_path), "case.pdb")
    open(path, "wb").write(bytes(g["pdb_text"]))
    pdb = PDB(path)
    pdb.set_coords(g["moved"])
    d = Dmap.from_array(G.copy(), v, origin=tuple(float(x) for x in O), normalize=False)
    rmsd, conv, step = refine_pdb(d, pdb, **kw)
    meta = g["refine_%s_meta" % tag]
    assert conv == bool(meta[1]) and step == int(meta[2])
    assert np.abs(pdb.coords - g["refine_%s_coords" % tag]).max() < 1e-9
    assert abs(rmsd - meta[0]) < 1e-9


@pytest.mark.gpu
def test_device_refine_batch_and_recovery():
    """Batched poses: every CTA reproduces the single-pose result; a displaced structure comes back onto its own
    density (size-independent property at a larger size than the fixture)."""
    torch = _cuda()
    import synth
    from mad_b200.Dmap import Dmap
    from mad_b200.structure_utils import RefineField, refine_poses
    from mad_b200.math_utils import euler_rod_mat
    g, G, O, v, *_ = _case()
    d = Dmap.from_array(G.copy(), v, origin=tuple(float(x) for x in O), normalize=False)
    field = RefineField(d)
    poses = np.stack([g["moved"], g["moved"] + np.array([0.4, 0.0, -0.3]), g["moved"]])
    coords, conv, step, bad = refine_poses(field, poses, n_steps=500, max_step_size=1, min_step_size=0.1)
    assert not bad.any() and conv.all()
    assert np.array_equal(coords[0], coords[2]) and step[0] == step[2]
    assert np.abs(coords[0] - g["refine_mad_coords"]).max() < 1e-9
    # recovery on a 20000-atom structure and its own simulated map
    atoms = synth.random_walk_atoms(20000, 120.0, 5)
    grid, org = synth.simulate_density(atoms, 8.0, 2.0)
    big = Dmap.from_array(np.asarray(grid, dtype=np.float32), 2.0, origin=tuple(float(x) for x in org), normalize=False)
    cen = atoms.mean(0)
    moved = np.dot(atoms - cen, euler_rod_mat(np.array([0.0, 0.6, 0.8]), 0.02)) + cen + np.array([1.5, -1.0, 0.5])
    out, conv, step, bad = refine_poses(RefineField(big), moved, n_steps=500, max_step_size=1, min_step_size=0.05)
    before = np.sqrt(np.mean(np.sum((moved - atoms) ** 2, axis=1)))
    after = np.sqrt(np.mean(np.sum((out[0] - atoms) ** 2, axis=1)))
    assert conv[0] and after < 0.25 * before, (before, after)


def test_host_common_box_equals_oracle_bounds():
    """The host-side bounds of Dmap._common_box (what mad_box_scores is launched with) against the oracle's restatement
    of mad/Dmap.py:172-241 on random integer and half-integer offsets (CPU only)."""
    import score_oracle as so
    from mad_b200.Dmap import Dmap
    rng = np.random.default_rng(3)
    n_checked = 0
    for _ in range(1200):
        s1 = [int(v) for v in rng.integers(3, 40, 3)]
        s2 = [int(v) for v in rng.integers(3, 40, 3)]
        o1 = [float(v) / 2 for v in rng.integers(-30, 30, 3)]       # integer and half-integer voxel offsets
        o2 = [float(v) / 2 for v in rng.integers(-30, 30, 3)]
        want = so.common_box(o1, s1, o2, s2)
        try:
            got = Dmap._common_box(o1, s1, o2, s2)
        except ValueError:
            got = "mismatch"
        if want is None:
            assert got is None
            continue
        g1 = np.zeros(s1)[tuple(slice(b[0], b[2]) for b in want)]
        g2 = np.zeros(s2)[tuple(slice(b[1], b[3]) for b in want)]
        if g1.shape != g2.shape:                       # the reference's np.dot would raise on such boxes
            assert got == "mismatch"
            continue
        assert got[:3] == [b[0] for b in want] and got[3:6] == [b[1] for b in want] and got[6:] == list(g1.shape)
        n_checked += 1
    assert n_checked > 60


def test_pdb_helpers_equal_reference(tmp_path):
    """PDB.rotate_atoms / translate_atoms / write_pdb / rmsd (host glue, mad/PDB.py:80-128) against the reference's output."""
    from mad_b200.PDB import PDB
    from mad_b200.math_utils import euler_rod_mat
    g = H.golden("score")
    path = os.path.join(str(tmp_path), "case.pdb")
    open(path, "wb").write(bytes(g["pdb_text"]))
    pa, pb = PDB(path), PDB(path)
    pb.set_coords(g["moved"])
    pb.rotate_atoms(euler_rod_mat([0, 0, 1], 0.3))
    pb.translate_atoms([1.25, -2.5, 3.75])
    out = os.path.join(str(tmp_path), "written.pdb")
    pb.write_pdb(out)
    assert open(out, "rb").read() == bytes(g["pdb_written"])
    assert np.allclose([pa.get_rmsd_with(pb), pa.get_rmsdCA_with(pb)], g["pdb_rmsd"], rtol=1e-14, atol=0)
