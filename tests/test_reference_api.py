"""The reference-shaped classes (MapSpace / Detector / Orientator / Descriptor / DensityFeature, the drop-in
boundary of mad/MaD.py:358-368) reproduce the reference's feature lists from an MRC file."""
import copy
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_describe_struct_through_the_reference_api(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import synth
    from mad_b200 import mrc
    from mad_b200.MapSpace import MapSpace
    from mad_b200.Detector import Detector
    from mad_b200.Orientator import Orientator
    from mad_b200.Descriptor import Descriptor

    g = H.golden("small")
    grid = synth.dequantise_u16(g["input_q"])
    v = float(g["voxelsp"])
    path = os.path.join(str(tmp_path), "small.mrc")
    mrc.write_mrc(path, grid.transpose(2, 1, 0), v, origin=tuple(g["origin"]))

    # the order of mad/MaD.py:358-368
    ms = MapSpace(path, resolution=8.0, voxelsp=0, sig_init=2)
    ms.build_space()
    assert np.allclose([ms.xi, ms.yi, ms.zi], g["ms_origin"]) and np.allclose(ms.voxelsp_list, g["voxelsp_list"])
    assert H.sha_flushed(ms.map_space[1]) == str(g["log1_sha256_flushed"])
    assert H.sha_flushed(np.ascontiguousarray(ms.grad_list[0])) == str(g["grad0_sha256_flushed"])
    anchors = Detector().find_anchors(ms)
    assert len(anchors) == len(g["kp_oct"])
    assert [a.index for a in anchors] == list(g["kp_index"])
    assert np.array_equal([a.oct_scale for a in anchors], g["kp_oct"])
    assert np.array_equal(np.array([a.coords for a in anchors]), g["kp_coords"])
    assert np.array_equal(np.array([a.map_coords for a in anchors]), g["kp_map_coords"])
    assert np.abs(np.array([a.subv_map_coords for a in anchors]) - g["kp_subv_map_coords"]).max() <= 1e-6
    oriented = Orientator(ori_radius=16).assign_orientations(ms, anchors)
    described = Descriptor(dsc_radius=16).generate_descriptors(ms, oriented)
    assert len(described) == len(g["of_index"])
    assert np.array_equal([d.index for d in described], g["of_index"])
    assert np.array_equal([d.oct_scale for d in described], g["of_oct"])
    assert np.array_equal([d.main_bin for d in described], g["of_main"])
    assert np.array_equal([d.sec_bin for d in described], g["of_sec"])
    assert np.array_equal(np.array([d.coords for d in described]), g["of_coords"])
    assert np.abs(np.array([d.subv_map_coords for d in described]) - g["of_subv_map_coords"]).max() <= 1e-6
    assert np.array_equal(np.array([d.lin_ar_subeqsp for d in described]), g["dsc"])
    assert all(d.eqsp_size == 112 and d.subeqsp_size == 16 for d in described)
    rf = {(int(a), int(b)): m for (a, b), m in zip(g["rfinal_ab"], g["rfinal_mat"])}
    for d in described:
        assert np.abs(np.asarray(d.Rfinal) - rf[(d.main_bin, d.sec_bin)]).max() <= 1e-12

    # feature lists that do not come straight from this package (plain Python lists, e.g. rebuilt from
    # the reference's HDF5 cache): the stages rebuild the device tables from the attributes
    plain = [copy.copy(a) for a in anchors]
    oriented2 = Orientator(ori_radius=16).assign_orientations(ms, plain)
    assert [(d.index, d.main_bin, d.sec_bin) for d in oriented2] == [(d.index, d.main_bin, d.sec_bin) for d in described]
    described2 = Descriptor(dsc_radius=16).generate_descriptors(ms, [copy.copy(d) for d in oriented2])
    assert np.array_equal(np.array([d.lin_ar_subeqsp for d in described2]), g["dsc"])


def test_argument_errors_follow_the_reference(tmp_path, capsys):
    from mad_b200.MapSpace import MapSpace
    with pytest.raises(SystemExit):
        MapSpace("model.pdb", resolution=0, voxelsp=2.0)             # mad/MapSpace.py:57-62
    with pytest.raises(SystemExit):
        MapSpace("something.txt")                                    # mad/MapSpace.py:63-67
    assert "MaD> ERROR" in capsys.readouterr().out


def test_mapspace_pdb_mode_runs_on_the_device(tmp_path):
    """MapSpace(<pdb>, resolution, voxelsp) (mad/MapSpace.py:73-76): atoms -> density -> scale space without the
    reference's Python; the density equals the reference's simulation to one float32 ulp (tests/test_density.py)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mad_b200.MapSpace import MapSpace
    from mad_b200.Detector import Detector
    g = H.golden("density")
    path = os.path.join(str(tmp_path), "case.pdb")
    open(path, "wb").write(bytes(g["pdb_text"]))
    ms = MapSpace(path, resolution=8.0, voxelsp=2.0)
    ms.build_space()
    ref = g["a_grid"]
    assert tuple(ms.space.dims[1]) == tuple(s + 18 for s in ref.shape)
    assert np.allclose([ms.xi, ms.yi, ms.zi], g["a_origin"] - 9 * 2.0)
    base = ms.grid_list[1][9:-9, 9:-9, 9:-9]
    assert np.abs(base - ref).max() <= 1.2e-7
    assert len(Detector().find_anchors(ms)) > 0


@pytest.mark.parametrize("mode", ["up", "base"])
def test_single_octave_modes_equal_reference(tmp_path, mode):
    """MapSpace(oct_mode="up" / "base") (mad/MapSpace.py:149-163) through the reference-shaped classes against the
    reference's own run: one grid, octave index 0, stride-2 patch geometry."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import synth
    from mad_b200 import mrc
    from mad_b200.MapSpace import MapSpace
    from mad_b200.Detector import Detector
    from mad_b200.Orientator import Orientator
    from mad_b200.Descriptor import Descriptor
    g, gs = H.golden("octmode"), H.golden("small")
    grid = synth.dequantise_u16(gs["input_q"])
    path = os.path.join(str(tmp_path), "small.mrc")
    mrc.write_mrc(path, grid.transpose(2, 1, 0), float(gs["voxelsp"]), origin=tuple(gs["origin"]))
    ms = MapSpace(path, oct_mode=mode)
    ms.build_space()
    assert len(ms.map_space) == len(ms.grad_list) == len(ms.voxelsp_list) == 1
    assert np.allclose(ms.voxelsp_list, g[mode + "_voxelsp_list"])
    assert H.sha_flushed(ms.map_space[0]) == str(g[mode + "_log_sha256_flushed"])
    anchors = Detector().find_anchors(ms)
    assert np.array_equal(np.array([a.coords for a in anchors]).reshape(-1, 3), g[mode + "_kp_coords"])
    assert all(a.oct_scale == 0 for a in anchors)
    assert np.abs(np.array([a.subv_map_coords for a in anchors]).reshape(-1, 3) - g[mode + "_kp_subv_map_coords"]).max() <= 1e-6
    described = Descriptor(dsc_radius=16).generate_descriptors(ms, Orientator(ori_radius=16).assign_orientations(ms, anchors))
    assert np.array_equal([d.index for d in described], g[mode + "_of_index"])
    assert np.array_equal([d.main_bin for d in described], g[mode + "_of_main"]) and np.array_equal([d.sec_bin for d in described], g[mode + "_of_sec"])
    d = np.array([f.lin_ar_subeqsp for f in described], dtype=np.int16).reshape(-1, 1024)
    assert np.array_equal(H.crc_rows(d), g[mode + "_dsc_crc32"])
