"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  It imports the reference package
``mad`` through the shims of ``ref_shims.py`` from a scratch directory that holds a
``mad -> /root/reference/mad`` symlink (``mad/eqsp/eqsp.py:16,26`` opens its tables relative
to the cwd), runs ``MapSpace -> Detector -> Orientator -> Descriptor`` (the order of
``mad/MaD.py:358-368``) and ``MaD._match_dsc`` (``mad/MaD.py:414-453``) on seeded synthetic
inputs, and stores

* the input grid as uint16 levels (``grid = q / 65535`` in f32 -- exact and portable),
* SHA-256 + a few thousand sampled values of every dense intermediate
  (``grid_list[0]``, ``map_space``, ``gauss_list``, ``grad_list``),
* the full sparse results: keypoints, oriented (index, main, sec) triples, ``Rfinal``,
  descriptors (or their per-row CRC32 when large), pair lists and scores.

    python oracle/gen_goldens.py [tiny small pair c1]

The committed fixtures pin ``oracle/mad_oracle.py`` (tests/test_oracle_golden.py).
"""
import hashlib
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
WORK = "/tmp/mad_ref_work"
GOLD = os.path.join(REPO, "tests", "golden")

sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import synth  # noqa: E402


def _enter_workdir():
    os.makedirs(WORK, exist_ok=True)
    link = os.path.join(WORK, "mad")
    if not os.path.islink(link):
        os.symlink(os.path.join(REF, "mad"), link)
    os.chdir(WORK)
    sys.path.insert(0, WORK)
    ref_shims.install()


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


FLUSH = 1e-10


def sha_flushed(a):
    """SHA-256 with |v| < FLUSH set to +0: the spline's decaying tails in the zero padding
    (|v| ~ 1e-20) depend on the banded solver's rounding and are not part of the parity contract;
    everything above FLUSH must still be bit-identical."""
    a = np.ascontiguousarray(a).copy()
    a[np.abs(a) < FLUSH] = 0
    return hashlib.sha256(a.tobytes()).hexdigest()


def sample_positions(shape, n, seed):
    rng = np.random.default_rng(seed)
    return np.stack([rng.integers(0, s, size=n) for s in shape[:3]], axis=1).astype(np.int32)


def dense_record(out, key, arr, seed):
    arr = np.ascontiguousarray(arr)
    out[key + "_sha256"] = np.array(sha(arr))
    out[key + "_sha256_flushed"] = np.array(sha_flushed(arr))
    out[key + "_shape"] = np.array(arr.shape, dtype=np.int64)
    out[key + "_dtype"] = np.array(str(arr.dtype))
    pos = sample_positions(arr.shape, 4096, seed)
    out[key + "_pos"] = pos
    out[key + "_val"] = arr[pos[:, 0], pos[:, 1], pos[:, 2]]
    out[key + "_absmax"] = np.array(float(np.abs(arr).max()))


def run_pipeline(mrc_path, patch_size=16, want_hist=False, oct_mode="both"):
    from mad.MapSpace import MapSpace
    from mad.Detector import Detector
    from mad.Orientator import Orientator
    from mad.Descriptor import Descriptor
    Orientator.step1_reject = 0   # latent AttributeError, mad/Orientator.py:133,153

    t = {}
    ms = MapSpace(mrc_path, oct_mode=oct_mode)
    t0 = time.perf_counter(); ms.build_space(); t["build_space"] = time.perf_counter() - t0
    t0 = time.perf_counter(); anchors = Detector().find_anchors(ms); t["find_anchors"] = time.perf_counter() - t0
    ori = Orientator(ori_radius=patch_size)
    t0 = time.perf_counter(); oriented = ori.assign_orientations(ms, anchors); t["assign_orientations"] = time.perf_counter() - t0
    dsc = Descriptor(dsc_radius=patch_size)
    t0 = time.perf_counter(); described = dsc.generate_descriptors(ms, oriented); t["generate_descriptors"] = time.perf_counter() - t0
    return ms, anchors, described, t


def collect_case(q, voxelsp, origin, ms, anchors, described, timings, full_dsc=True):
    """Everything a fixture stores about one reference run, as a dict of arrays (+ the descriptors)."""
    out = {}
    if q is not None:
        out["input_q"] = q
    out["voxelsp"] = np.array(voxelsp, dtype=np.float64)
    out["origin"] = np.array(origin, dtype=np.float64)
    out["ms_origin"] = np.array([ms.xi, ms.yi, ms.zi], dtype=np.float64)
    out["voxelsp_list"] = np.array(ms.voxelsp_list, dtype=np.float64)
    dense_record(out, "up_grid", ms.grid_list[0], 11)
    for o in range(2):
        dense_record(out, "log%d" % o, ms.map_space[o], 20 + o)
        dense_record(out, "gauss%d" % o, ms.gauss_list[o], 30 + o)
        g = ms.grad_list[o]
        out["grad%d_sha256" % o] = np.array(sha(g))
        out["grad%d_sha256_flushed" % o] = np.array(sha_flushed(g))
        out["grad%d_shape" % o] = np.array(g.shape, dtype=np.int64)
        out["grad%d_dtype" % o] = np.array(str(g.dtype))
        pos = sample_positions(g.shape, 4096, 40 + o)
        out["grad%d_pos" % o] = pos
        out["grad%d_val" % o] = g[pos[:, 0], pos[:, 1], pos[:, 2], :]
    # keypoints (Detector.find_anchors output, in order)
    out["kp_index"] = np.array([a.index for a in anchors], dtype=np.int32)
    out["kp_oct"] = np.array([a.oct_scale for a in anchors], dtype=np.int32)
    out["kp_coords"] = np.array([a.coords for a in anchors], dtype=np.int32).reshape(-1, 3)
    out["kp_map_coords"] = np.array([a.map_coords for a in anchors], dtype=np.float64).reshape(-1, 3)
    out["kp_subv_map_coords"] = np.array([a.subv_map_coords for a in anchors], dtype=np.float64).reshape(-1, 3)
    out["kp_val"] = np.array([a.voxel_val for a in anchors], dtype=np.float32)
    out["kp_coord_dtypes"] = np.array("%s %s" % (np.asarray(anchors[0].map_coords).dtype,
                                                  np.asarray(anchors[0].subv_map_coords).dtype) if anchors else "")
    # oriented + described features (in order)
    out["of_index"] = np.array([d.index for d in described], dtype=np.int32)
    out["of_oct"] = np.array([d.oct_scale for d in described], dtype=np.int32)
    out["of_main"] = np.array([d.main_bin for d in described], dtype=np.int32)
    out["of_sec"] = np.array([d.sec_bin for d in described], dtype=np.int32)
    out["of_coords"] = np.array([d.coords for d in described], dtype=np.int32).reshape(-1, 3)
    out["of_subv_map_coords"] = np.array([d.subv_map_coords for d in described], dtype=np.float64).reshape(-1, 3)
    rf = np.array([d.Rfinal for d in described], dtype=np.float64).reshape(-1, 3, 3)
    # Rfinal depends only on (main, sec): store the unique table
    ab = out["of_main"].astype(np.int64) * 1000 + out["of_sec"]
    uab, first = np.unique(ab, return_index=True)
    out["rfinal_ab"] = np.stack([uab // 1000, uab % 1000], 1).astype(np.int32)
    out["rfinal_mat"] = rf[first]
    same = all(np.array_equal(rf[i], rf[first[np.searchsorted(uab, ab[i])]]) for i in range(len(ab)))
    out["rfinal_depends_only_on_ab"] = np.array(bool(same))
    dsc = np.array([d.lin_ar_subeqsp for d in described], dtype=np.int16).reshape(-1, 1024)
    out["dsc_sha256"] = np.array(sha(dsc))
    out["dsc_crc32"] = np.array([zlib.crc32(row.tobytes()) for row in dsc], dtype=np.uint32)
    out["dsc_rowsum"] = dsc.sum(1).astype(np.int32)
    if full_dsc:
        out["dsc"] = dsc
    out["ref_timings_s"] = np.array([timings[k] for k in ("build_space", "find_anchors",
                                                         "assign_orientations", "generate_descriptors")])
    return out, dsc


def pack_case(name, q, voxelsp, origin, ms, anchors, described, timings, full_dsc=True):
    out, dsc = collect_case(q, voxelsp, origin, ms, anchors, described, timings, full_dsc)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s  K=%d D=%d  (%.1f KB)  timings=%s" % (path, len(anchors), len(described),
                                                       os.path.getsize(path) / 1024, timings))
    return dsc


def add_flushed_digests(name):
    """Adds the *_sha256_flushed keys to an existing fixture by re-running only
    MapSpace.build_space of the reference on the stored input (sparse results are kept)."""
    from mad.MapSpace import MapSpace
    path = os.path.join(GOLD, name + ".npz")
    with np.load(path, allow_pickle=False) as z:
        out = {k: z[k] for k in z.files}
    grid = synth.dequantise_u16(out["input_q"])
    mrc_path = os.path.join(WORK, name + "_flush.mrc")
    ref_shims.write_mrc_stub(mrc_path, grid, float(out["voxelsp"]), tuple(out["origin"]))
    ms = MapSpace(mrc_path)
    ms.build_space()
    arrays = {"up_grid": ms.grid_list[0]}
    for o in range(2):
        arrays["log%d" % o] = ms.map_space[o]
        arrays["gauss%d" % o] = ms.gauss_list[o]
        arrays["grad%d" % o] = ms.grad_list[o]
    for k, a in arrays.items():
        assert sha(a) == str(out[k + "_sha256"]), "reference output changed for %s/%s" % (name, k)
        out[k + "_sha256_flushed"] = np.array(sha_flushed(a))
    np.savez_compressed(path, **out)
    print("updated %s with flushed digests" % path)


def density_from_atoms(coords, resolution, voxelsp, tag):
    """The reference's own simulator (mad/PDB.py:131), then uint16 quantisation."""
    from mad.PDB import PDB
    pdb_path = os.path.join(WORK, tag + ".pdb")
    synth.write_pdb(pdb_path, coords)
    grid, xi, yi, zi = PDB(pdb_path).structure_to_density(resolution, voxelsp)
    q = synth.quantise_u16(grid)
    return q, (xi, yi, zi)


def case_from_atoms(name, coords, resolution, voxelsp, full_dsc=True):
    q, origin = density_from_atoms(coords, resolution, voxelsp, name)
    origin = tuple(float(int(o)) for o in origin)          # int-truncated origin (MapSpace.py:108)
    grid = synth.dequantise_u16(q)
    mrc_path = os.path.join(WORK, name + ".mrc")
    ref_shims.write_mrc_stub(mrc_path, grid, voxelsp, origin)
    ms, anchors, described, t = run_pipeline(mrc_path)
    dsc = pack_case(name, q, voxelsp, origin, ms, anchors, described, t, full_dsc=full_dsc)
    return described, dsc


def rigid(coords, seed, shift):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(3, 3))
    qm, _ = np.linalg.qr(a)
    if np.linalg.det(qm) < 0:
        qm[:, 0] = -qm[:, 0]
    c = coords.mean(0)
    return (coords - c) @ qm.T + c + np.asarray(shift)


def main(which):
    _enter_workdir()
    os.makedirs(GOLD, exist_ok=True)
    if which and which[0] == "--add-flushed":
        for name in which[1:]:
            add_flushed_digests(name)
        return
    if "tiny" in which:
        case_from_atoms("tiny", synth.random_walk_atoms(400, 30.0, 5), 8.0, 2.0)
    if "small" in which:
        case_from_atoms("small", synth.random_walk_atoms(1500, 60.0, 4), 8.0, 2.0)
    if "pair" in which:
        # matching golden: a 2-subunit assembly map (lo) against one of its subunits (hi)
        from mad.MaD import MaD
        sub_a = synth.random_walk_atoms(1200, 50.0, 21)
        sub_b = synth.random_walk_atoms(1200, 50.0, 22)
        asm = np.concatenate([rigid(sub_a, 31, (70.0, 10.0, 5.0)), rigid(sub_b, 32, (10.0, 60.0, 40.0))])
        hi_list, hi_dsc = case_from_atoms("pair_hi", sub_a, 8.0, 2.0)
        lo_list, lo_dsc = case_from_atoms("pair_lo", asm, 8.0, 2.0)
        t0 = time.perf_counter()
        results, lo_cloud, hi_cloud = MaD()._match_dsc(lo_list, hi_list, cc_threshold=0.6)
        dt = time.perf_counter() - t0
        res = np.array(results, dtype=np.float64).reshape(-1, 23)
        # re-derive the raw contraction outputs exactly as mad/MaD.py:416-424 does
        def unit_rows(dl):
            return np.array([d.lin_ar_subeqsp / np.linalg.norm(d.lin_ar_subeqsp)
                             if np.linalg.norm(d.lin_ar_subeqsp) > 0 else d.lin_ar_subeqsp for d in dl])
        preds = np.dot(unit_rows(hi_list), unit_rows(lo_list).T)
        pairs = np.array(np.where(preds > 0.6)).T.astype(np.int32)
        out = dict(pairs=pairs, scores=preds[pairs[:, 0], pairs[:, 1]], cc=np.array(0.6),
                   results=res, lo_cloud=lo_cloud, hi_cloud=hi_cloud,
                   preds_sha256=np.array(sha(preds)), preds_shape=np.array(preds.shape),
                   topk8_idx=np.argsort(-preds, axis=1, kind="stable")[:, :8].astype(np.int32),
                   match_time_s=np.array(dt))
        assert np.array_equal(res[:, 0], out["scores"])
        path = os.path.join(GOLD, "pair_match.npz")
        np.savez_compressed(path, **out)
        print("wrote %s  M=%d N=%d pairs=%d" % (path, preds.shape[0], preds.shape[1], len(pairs)))
    if "dmap" in which:
        # Dmap container (row a0): the reference's own class on a Situs text file
        from mad.Dmap import Dmap
        rng = np.random.default_rng(9)
        shape = (20, 22, 24)
        x, y, z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
        g = np.zeros(shape, dtype=np.float64)
        for _ in range(5):
            p = np.array([10, 11, 12]) + rng.integers(-4, 5, size=3)
            g += np.exp(-((x - p[0]) ** 2 + (y - p[1]) ** 2 + (z - p[2]) ** 2) / 5.0)
        g -= 0.02                                            # some negative background
        g[g < 1e-3] *= (np.abs(g[g < 1e-3]) > 5e-3)          # exact zeros around the blob
        path = os.path.join(WORK, "dmap_case.sit")
        with open(path, "w") as f:
            f.write("%f %f %f %f %i %i %i\n\n" % (1.5, -3.0, 4.5, 6.0, shape[0], shape[1], shape[2]))
            k = 0
            for zz in range(shape[2]):
                for yy in range(shape[1]):
                    for xx in range(shape[0]):
                        f.write("   %6.6f " % g[xx][yy][zz])
                        k += 1
                        if k % 10 == 0:
                            f.write("\n")
        out = {"sit_text": np.frombuffer(open(path, "rb").read(), dtype=np.uint8)}
        for tag, kw in (("a", dict(isovalue=0.3)), ("b", dict(isovalue=0.0, normalize=False, pad=3)), ("c", dict(isovalue=50.0))):
            d = Dmap(path, **kw)
            out["%s_ctor" % tag] = np.array(d.grid3d, dtype=np.float32)
            out["%s_ctor_meta" % tag] = np.array([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], dtype=np.float64)
            d.reduce_void()
            out["%s_void" % tag] = np.array(d.grid3d, dtype=np.float32)
            out["%s_void_meta" % tag] = np.array([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], dtype=np.float64)
            d.pad_grid(2)
            out["%s_pad_meta" % tag] = np.array([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], dtype=np.float64)
        np.savez_compressed(os.path.join(GOLD, "dmap.npz"), **out)
        print("wrote dmap.npz", {k: v.shape for k, v in out.items()})
    if "density" in which:
        # atoms -> density: the reference's own PDB.structure_to_density
        from mad.PDB import PDB
        coords = synth.random_walk_atoms(300, 30.0, 77)
        pdb_path = os.path.join(WORK, "density_case.pdb")
        synth.write_pdb(pdb_path, coords)
        lines = open(pdb_path).read().splitlines()
        for i in range(0, len(lines), 7):                       # a few nitrogens / oxygens / one unknown element
            lines[i] = lines[i][:76] + (" N" if i % 14 else " O")
        lines[5] = lines[5][:76] + "XX"
        open(pdb_path, "w").write("\n".join(lines) + "\n")
        out = {"pdb_text": np.frombuffer(open(pdb_path, "rb").read(), dtype=np.uint8)}
        for tag, kw in (("a", dict(resolution=8.0, voxelsp=2.0)), ("b", dict(resolution=4.0, voxelsp=1.0, isovalue=0.2)),
                        ("c", dict(resolution=5.0, voxelsp=2.0, isovalue=0.2, pad=1))):
            g, dxi, dyi, dzi = PDB(pdb_path).structure_to_density(**kw)
            out[tag + "_grid"] = g
            out[tag + "_origin"] = np.array([dxi, dyi, dzi], dtype=np.float64)
            if tag == "c":                                      # the reference's Situs writer (mad/PDB.py:165-179)
                sit_path = os.path.join(WORK, "density_case_c.sit")
                PDB(pdb_path).structure_to_density(outname=sit_path, **kw)
                out["c_sit_text"] = np.frombuffer(open(sit_path, "rb").read(), dtype=np.uint8)
        np.savez_compressed(os.path.join(GOLD, "density.npz"), **out)
        print("wrote density.npz", {k: v.shape for k, v in out.items()})
    if "score" in which:
        # scoring / refinement numerics: the reference's own Dmap.get_CCC_with_grid,
        # get_CCC_with_dmap, mask_with, structure_utils.get_overlap and structure_utils.refine_pdb
        from mad.PDB import PDB
        from mad.Dmap import Dmap
        from mad.structure_utils import refine_pdb, get_overlap
        from mad.math_utils import euler_rod_mat

        def bare_dmap(grid, voxsp, origin):
            d = Dmap.__new__(Dmap)
            d.voxsp = voxsp
            d.xi, d.yi, d.zi = [float(v) for v in origin]
            d.grid3d = np.array(grid, dtype=np.float32)
            d.xb, d.yb, d.zb = d.grid3d.shape
            d.map_name = d.name = "bare"
            return d

        coords = synth.random_walk_atoms(600, 40.0, 41)
        pdb_path = os.path.join(WORK, "score_case.pdb")
        synth.write_pdb(pdb_path, coords)
        coords = PDB(pdb_path).get_coords().copy()               # as parsed (%8.3f columns)
        g_map, mxi, myi, mzi = PDB(pdb_path).structure_to_density(8.0, 2.0)
        g_map = np.pad(g_map, 4)
        m_org = np.array([mxi - 8.0, myi - 8.0, mzi - 8.0])
        out = {"pdb_text": np.frombuffer(open(pdb_path, "rb").read(), dtype=np.uint8),
               "map_grid": np.array(g_map, dtype=np.float32), "map_origin": m_org, "voxsp": np.array(2.0)}
        # a displaced pose of the same structure and its simulated density
        cen = coords.mean(0)
        moved = np.dot(coords - cen, euler_rod_mat(np.array([0.3, -0.5, 0.81]) / np.linalg.norm([0.3, -0.5, 0.81]), 0.12)) \
            + cen + np.array([2.3, -1.7, 1.1])
        out["moved"] = moved
        for tag, kw in (("mad", dict(n_steps=500, max_step_size=1, min_step_size=0.1)), ("default", dict()),
                        ("short", dict(n_steps=7, max_step_size=1, min_step_size=0.1))):
            pdb = PDB(pdb_path)
            pdb.set_coords(moved)
            t0 = time.perf_counter()
            rmsd, conv, step = refine_pdb(bare_dmap(g_map, 2.0, m_org), pdb, **kw)
            dt = time.perf_counter() - t0
            out["refine_%s_coords" % tag] = pdb.coords.copy()
            out["refine_%s_meta" % tag] = np.array([rmsd, float(conv), float(step), dt])
            print("refine", tag, rmsd, conv, step, "%.2fs" % dt,
                  "rmsd to truth %.4f" % np.sqrt(np.mean(np.sum((pdb.coords - coords) ** 2, axis=1))))
        pdb = PDB(pdb_path)
        pdb.set_coords(moved)
        g_sub, sxi, syi, szi = pdb.structure_to_density(8.0, 2.0)
        out["sub_grid"] = np.array(g_sub, dtype=np.float32)
        out["sub_origin"] = np.array([sxi, syi, szi])
        for tag, iso in (("iso0", 0), ("iso2", 0.2)):
            d = bare_dmap(g_map, 2.0, m_org)
            out["ccc_grid_" + tag] = np.array(d.get_CCC_with_grid(g_sub.copy(), sxi, syi, szi, isovalue=iso), dtype=np.float64)
            d = bare_dmap(g_map, 2.0, m_org)
            out["ccc_dmap_" + tag] = np.array(d.get_CCC_with_dmap(bare_dmap(g_sub, 2.0, (sxi, syi, szi)), isovalue=iso),
                                              dtype=np.float64)
        # a sub grid sticking out of the map on the low side of x and the high side of z, and a disjoint one
        far = np.array([sxi - 30.0, syi + 4.0, szi + 26.0])
        out["far_origin"] = far
        out["ccc_grid_far"] = np.array(bare_dmap(g_map, 2.0, m_org).get_CCC_with_grid(g_sub.copy(), *far), dtype=np.float64)
        out["ccc_dmap_far"] = np.array(bare_dmap(g_map, 2.0, m_org).get_CCC_with_dmap(bare_dmap(g_sub, 2.0, far)),
                                       dtype=np.float64)
        out["overlap"] = np.array(get_overlap([g_map.copy(), *m_org], [g_sub.copy(), sxi, syi, szi], 2), dtype=np.float64)
        out["overlap_far"] = np.array(get_overlap([g_map.copy(), *m_org], [g_sub.copy(), *far], 2), dtype=np.float64)
        out["overlap_iso"] = np.array(get_overlap([g_map.copy(), *m_org], [g_sub.copy(), sxi, syi, szi], 2, isovalue=0.3),
                                      dtype=np.float64)
        d = bare_dmap(g_map, 2.0, m_org)
        d.mask_with(bare_dmap(g_sub, 2.0, (sxi, syi, szi)))
        out["masked"] = d.grid3d.copy()
        d = bare_dmap(g_map, 2.0, m_org)
        d.mask_with(bare_dmap(g_sub, 2.0, far))
        out["masked_far"] = d.grid3d.copy()
        # PDB manipulation helpers used around the refinement (mad/PDB.py:80-128): written file and RMSDs
        pa, pb = PDB(pdb_path), PDB(pdb_path)
        pb.set_coords(moved)
        pb.rotate_atoms(euler_rod_mat([0, 0, 1], 0.3))
        pb.translate_atoms([1.25, -2.5, 3.75])
        wpath = os.path.join(WORK, "score_written.pdb")
        pb.write_pdb(wpath)
        out["pdb_written"] = np.frombuffer(open(wpath, "rb").read(), dtype=np.uint8)
        out["pdb_rmsd"] = np.array([pa.get_rmsd_with(pb), pa.get_rmsdCA_with(pb)], dtype=np.float64)
        np.savez_compressed(os.path.join(GOLD, "score.npz"), **out)
        print("wrote score.npz", {k: (v.shape if v.ndim else float(v)) for k, v in out.items() if k != "pdb_text"})
    if "octmode" in which:
        # oct_mode "up" / "base" (mad/MapSpace.py:149-163; never selected by MaD.run): sparse results on the `small` input
        with np.load(os.path.join(GOLD, "small.npz"), allow_pickle=False) as z:
            q, voxelsp, origin = z["input_q"], float(z["voxelsp"]), tuple(z["origin"])
        mrc_path = os.path.join(WORK, "octmode.mrc")
        ref_shims.write_mrc_stub(mrc_path, synth.dequantise_u16(q), voxelsp, origin)
        out = {}
        for mode in ("up", "base"):
            ms, anchors, described, t = run_pipeline(mrc_path, oct_mode=mode)
            out[mode + "_voxelsp_list"] = np.array(ms.voxelsp_list, dtype=np.float64)
            out[mode + "_n_grids"] = np.array(len(ms.map_space))
            out[mode + "_log_sha256_flushed"] = np.array(sha_flushed(ms.map_space[0]))
            out[mode + "_kp_oct"] = np.array([a.oct_scale for a in anchors], dtype=np.int32)
            out[mode + "_kp_coords"] = np.array([a.coords for a in anchors], dtype=np.int32).reshape(-1, 3)
            out[mode + "_kp_subv_map_coords"] = np.array([a.subv_map_coords for a in anchors], dtype=np.float64).reshape(-1, 3)
            out[mode + "_of_index"] = np.array([d.index for d in described], dtype=np.int32)
            out[mode + "_of_main"] = np.array([d.main_bin for d in described], dtype=np.int32)
            out[mode + "_of_sec"] = np.array([d.sec_bin for d in described], dtype=np.int32)
            dsc = np.array([d.lin_ar_subeqsp for d in described], dtype=np.int16).reshape(-1, 1024)
            out[mode + "_dsc_crc32"] = np.array([zlib.crc32(r.tobytes()) for r in dsc], dtype=np.uint32)
            print("octmode", mode, "K=%d D=%d" % (len(anchors), len(described)))
        np.savez_compressed(os.path.join(GOLD, "octmode.npz"), **out)
    if "c1" in which:
        case_from_atoms("c1", synth.random_walk_atoms(9000, 85.0, 1), 4.0, 1.0, full_dsc=False)


if __name__ == "__main__":
    main(sys.argv[1:] or ["tiny", "small", "pair"])
