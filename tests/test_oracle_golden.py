"""The oracle (oracle/mad_oracle.py) against the committed golden fixtures, which were produced by
running the UNMODIFIED reference (oracle/gen_goldens.py, build container only).  CPU only.

Dense arrays are pinned by SHA-256 (bit-exact), sparse results element by element.
"""
import numpy as np
import pytest

import helpers as H
import mad_oracle as mo
import synth


@pytest.mark.parametrize("case", ["tiny", "small", "pair_hi"])
def test_oracle_reproduces_reference_case(case):
    g = H.golden(case)
    grid, sp, kp, ori, dsc, tab_o = H.oracle_case(case)
    assert H.sha(sp["grid_list"][0]) == str(g["up_grid_sha256"])
    for o in range(2):
        assert H.sha(sp["map_space"][o]) == str(g["log%d_sha256" % o])
        assert H.sha(sp["gauss_list"][o]) == str(g["gauss%d_sha256" % o])
        assert H.sha(sp["grad_list"][o]) == str(g["grad%d_sha256" % o])
        assert sp["map_space"][o].dtype == np.float32 and sp["grad_list"][o].dtype == np.float32
    # keypoints in the reference's order
    assert np.array_equal(kp["oct"], g["kp_oct"])
    assert np.array_equal(kp["coords"], g["kp_coords"])
    assert np.array_equal(kp["val"], g["kp_val"])
    assert np.array_equal(kp["map_coords"], g["kp_map_coords"])
    assert np.array_equal(kp["subv_map_coords"], g["kp_subv_map_coords"])
    # oriented features: (index, main, sec) in emission order
    assert np.array_equal(kp["index"][ori["kp"]], g["of_index"])
    assert np.array_equal(ori["main"], g["of_main"])
    assert np.array_equal(ori["sec"], g["of_sec"])
    # Rfinal depends only on (main, sec)
    assert bool(g["rfinal_depends_only_on_ab"])
    for (a, b), m in zip(g["rfinal_ab"], g["rfinal_mat"]):
        assert np.array_equal(tab_o.rf(int(a), int(b)), m)
    # descriptors
    assert dsc.dtype == np.int16
    assert np.array_equal(dsc, g["dsc"])
    assert H.sha(dsc) == str(g["dsc_sha256"])


def test_oracle_c1_stencils_and_keypoints():
    """96^3 component map (BASELINE config 1): dense stages + detection in full, the per-keypoint
    stages on a slice (the whole case takes minutes on the CPU)."""
    g = H.golden("c1")
    import synth
    grid = synth.dequantise_u16(g["input_q"])
    sp = mo.build_space(grid)
    assert H.sha(sp["grid_list"][0]) == str(g["up_grid_sha256"])
    for o in range(2):
        assert H.sha(sp["map_space"][o]) == str(g["log%d_sha256" % o])
        assert H.sha(sp["grad_list"][o]) == str(g["grad%d_sha256" % o])
    v = float(g["voxelsp"])
    org = np.asarray(g["origin"], dtype=np.float64) - 9 * v
    kp = mo.detect(sp["map_space"], [v / 2, v], org)
    assert np.array_equal(kp["coords"], g["kp_coords"]) and np.array_equal(kp["oct"], g["kp_oct"])
    assert np.array_equal(kp["subv_map_coords"], g["kp_subv_map_coords"])
    # orient + describe for keypoints 0..39 and the last 40 (base octave)
    sel = np.r_[0:40, len(kp["oct"]) - 40:len(kp["oct"])]
    sub = {k: v_[sel] for k, v_ in kp.items()}
    ori, tab_o = mo.orient(sp["grad_list"], sub)
    rows = np.nonzero(np.isin(g["of_index"], sel))[0]
    assert np.array_equal(sel[ori["kp"]], g["of_index"][rows])
    assert np.array_equal(ori["main"], g["of_main"][rows]) and np.array_equal(ori["sec"], g["of_sec"][rows])
    dsc = mo.describe(sp["grad_list"], sub, ori, tab_o)
    assert np.array_equal(H.crc_rows(dsc), g["dsc_crc32"][rows])
    assert np.array_equal(dsc.sum(1), g["dsc_rowsum"][rows])


def test_oracle_matching_against_reference_pairs():
    ghi, glo, gm = H.golden("pair_hi"), H.golden("pair_lo"), H.golden("pair_match")
    pairs, scores = mo.match_threshold(ghi["dsc"], glo["dsc"], float(gm["cc"]))
    assert np.array_equal(pairs, gm["pairs"])
    assert np.array_equal(scores, gm["scores"])            # same NumPy, same BLAS: bit-equal
    preds = mo.match_scores(ghi["dsc"], glo["dsc"])
    assert H.sha(preds) == str(gm["preds_sha256"])
    idx, _ = mo.match_topk(ghi["dsc"], glo["dsc"], 8)
    assert np.array_equal(idx, gm["topk8_idx"])


def test_exact_integer_score_formula_matches_reference_scores():
    """The CUDA kernels score with dot / sqrt(n_a n_b) from exact integers."""
    ghi, glo, gm = H.golden("pair_hi"), H.golden("pair_lo"), H.golden("pair_match")
    hi, lo = ghi["dsc"].astype(np.int64), glo["dsc"].astype(np.int64)
    dot = hi @ lo.T
    n2a, n2b = (hi * hi).sum(1), (lo * lo).sum(1)
    sc = dot / np.sqrt((n2a[:, None] * n2b[None, :]).astype(np.float64))
    i, j = np.where(sc > 0.6)
    assert np.array_equal(np.stack([i, j], 1), gm["pairs"])
    assert np.abs(sc[i, j] - gm["scores"]).max() < 1e-15


def test_oracle_edge_cases():
    z = np.zeros((20, 22, 24), dtype=np.float32)
    sp = mo.build_space(z)
    assert sp["grid_list"][0].shape == (75, 79, 83)
    kp = mo.detect(sp["map_space"], [1.0, 2.0], [0, 0, 0])
    assert len(kp["oct"]) == 0
    ori, tab_o = mo.orient(sp["grad_list"], kp)
    assert len(ori["kp"]) == 0
    assert mo.describe(sp["grad_list"], kp, ori, tab_o).shape == (0, 1024)
    e = np.zeros((0, 1024), dtype=np.int16)
    p, s = mo.match_threshold(e, e)
    assert p.shape == (0, 2)
    # zero descriptors stay zero vectors: score 0 with everything (mad/MaD.py:416-417)
    a = np.zeros((2, 1024), dtype=np.int16)
    a[1, :4] = 3
    p, s = mo.match_threshold(a, a, 0.6)
    assert p.tolist() == [[1, 1]]


def _feature_dict(g):
    rf = {(int(a), int(b)): m for (a, b), m in zip(g["rfinal_ab"], g["rfinal_mat"])}
    R = np.array([rf[(int(a), int(b))] for a, b in zip(g["of_main"], g["of_sec"])])
    return dict(dsc=g["dsc"], subv=g["of_subv_map_coords"], index=g["of_index"], oct=g["of_oct"], main=g["of_main"],
                sec=g["of_sec"], Rfinal=R)


def test_oracle_match_dsc_loop_equals_reference():
    """mad/MaD.py:426-453: the per-pair repeatability loop, against the reference's own `results` table."""
    import mad_oracle as mo
    ghi, glo, gm = H.golden("pair_hi"), H.golden("pair_lo"), H.golden("pair_match")
    res, lo_cloud, hi_cloud = mo.match_dsc(_feature_dict(glo), _feature_dict(ghi), 4, float(gm["cc"]))
    assert np.array_equal(lo_cloud, gm["lo_cloud"]) and np.array_equal(hi_cloud, gm["hi_cloud"])
    assert res.shape == gm["results"].shape
    assert np.array_equal(res[:, 1:14], gm["results"][:, 1:14])            # repeatability, indices, coordinates
    assert np.abs(res[:, 0] - gm["results"][:, 0]).max() < 1e-14 and np.abs(res[:, 14:] - gm["results"][:, 14:]).max() < 1e-14


@pytest.mark.parametrize("mode", ["up", "base"])
def test_oracle_single_octave_modes_equal_reference(mode):
    """oct_mode "up" / "base" (mad/MapSpace.py:149-163): the lone grid is octave 0, so the later stages sample it with the
    stride-2 patch geometry of octave 0 -- keypoints, triples and descriptors against the reference's own run."""
    g, gs = H.golden("octmode"), H.golden("small")
    grid = synth.dequantise_u16(gs["input_q"])
    sp = mo.build_space(grid, oct_mode=mode)
    assert len(sp["map_space"]) == int(g[mode + "_n_grids"]) == 1
    assert H.sha_flushed(sp["map_space"][0]) == str(g[mode + "_log_sha256_flushed"])
    v = float(gs["voxelsp"])
    org = np.asarray(gs["origin"], dtype=np.float64) - 9 * v
    kp = mo.detect(sp["map_space"], list(g[mode + "_voxelsp_list"]), org)
    assert np.array_equal(kp["coords"], g[mode + "_kp_coords"]) and np.array_equal(kp["oct"], g[mode + "_kp_oct"])
    ori, tab = mo.orient(sp["grad_list"], kp)
    assert np.array_equal(ori["kp"], g[mode + "_of_index"]) and np.array_equal(ori["main"], g[mode + "_of_main"])
    assert np.array_equal(ori["sec"], g[mode + "_of_sec"])
    dsc = mo.describe(sp["grad_list"], kp, ori, tab)
    assert np.array_equal(H.crc_rows(dsc), g[mode + "_dsc_crc32"])


def test_oracle_port_equals_reference_on_a_benchmarked_snapshot():
    """The port that `bench.py --impl reference` times, on config C4's snapshot 63 (one of the maps bench.py runs): same
    keypoints, (index, main, sec) triples and descriptor rows as the UNMODIFIED reference (tests/golden/c4.npz)."""
    g = H.golden("c4")
    grid = synth.c4_snapshot(63)
    assert H.sha(np.ascontiguousarray(grid, dtype=np.float32)) == str(g["s63_input_sha256"])
    sp, kp, ori, dsc = mo.describe_struct(grid, 1.0)
    assert np.array_equal(kp["coords"], g["s63_kp_coords"]) and np.array_equal(kp["oct"], g["s63_kp_oct"])
    assert np.array_equal(ori["kp"], g["s63_of_index"]) and np.array_equal(ori["main"], g["s63_of_main"])
    assert np.array_equal(ori["sec"], g["s63_of_sec"])
    assert np.array_equal(H.crc_rows(dsc), g["s63_dsc_crc32"])
    for key, a in (("log0", sp["map_space"][0]), ("log1", sp["map_space"][1]), ("up_grid", sp["grid_list"][0])):
        assert H.sha_flushed(a) == str(g["s63_%s_sha256_flushed" % key]), key
