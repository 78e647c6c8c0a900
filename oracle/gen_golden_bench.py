"""TEST INFRASTRUCTURE ONLY -- fixtures for the BENCHMARKED configurations, from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; same shims / scratch directory as gen_goldens.py).
The inputs are the ones bench.py times (oracle/synth.py: c2_inputs, c4_snapshot, c5_descriptor_sets):

    python oracle/gen_golden_bench.py c2map          # C2 assembly map 256^3          -> tests/golden/c2.npz
    python oracle/gen_golden_bench.py c2planes       # + per-x-plane CRC32 of the dense arrays (re-runs build_space only)
    python oracle/gen_golden_bench.py c2comp 0 1 2   # C2 component maps (any subset)  -> /tmp parts
    python oracle/gen_golden_bench.py c2pairs        # merge parts + hi_all x lo pairs -> tests/golden/c2_comp.npz
    python oracle/gen_golden_bench.py c4 0 31 63     # three of the 64 C4 snapshots    -> tests/golden/c4.npz
    python oracle/gen_golden_bench.py c5             # 256 sampled hi rows x 100 000 lo, stable top-8 -> tests/golden/c5.npz

Each map goes MapSpace -> Detector -> Orientator -> Descriptor of the reference (mad/MaD.py:358-368); the pair list
is the reference's two lines (mad/MaD.py:420-424) evaluated in row blocks (the 43 643 x 40 230 float64 score matrix
is 14 GB).  Large tables are stored as per-row CRC32 + row sums, pair lists as count + per-hi-row counts + CRC32.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_goldens as G  # noqa: E402
import ref_shims  # noqa: E402
import synth  # noqa: E402

PARTS = "/tmp/mad_ref_work/parts"


def run_map(tag, grid, voxelsp):
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    mrc = os.path.join(G.WORK, tag + ".mrc")
    ref_shims.write_mrc_stub(mrc, grid, voxelsp, (0.0, 0.0, 0.0))
    ms, anchors, described, t = G.run_pipeline(mrc)
    out, dsc = G.collect_case(None, voxelsp, (0.0, 0.0, 0.0), ms, anchors, described, t, full_dsc=False)
    out["input_sha256"] = np.array(G.sha(grid))
    out["input_shape"] = np.array(grid.shape, dtype=np.int64)
    print("%s: K=%d D=%d timings=%s" % (tag, len(anchors), len(described), t), flush=True)
    return out, dsc


def plane_crcs(a):
    """CRC32 of every x plane of a dense array after flushing |v| < 1e-10 (tests/helpers.py:plane_crcs is the twin)."""
    a = np.ascontiguousarray(a)
    out = np.zeros(a.shape[0], dtype=np.uint32)
    for x in range(a.shape[0]):
        p = a[x].copy()
        p[np.abs(p) < G.FLUSH] = 0
        out[x] = zlib.crc32(p.tobytes())
    return out


def unit_rows(d):
    """mad/MaD.py:416-419 (zero rows stay zero)."""
    d = d.astype(np.float64)
    n = np.linalg.norm(d, axis=1)
    out = d.copy()
    nz = n > 0
    out[nz] = d[nz] / n[nz][:, None]
    return out


def main(argv):
    G._enter_workdir()
    os.makedirs(PARTS, exist_ok=True)
    what = argv[0]
    if what == "c2map":
        grid, _ = synth.c2_inputs(0)
        out, dsc = run_map("c2map", grid, synth.C2["voxelsp"])
        np.save(os.path.join(PARTS, "c2map_dsc.npy"), dsc)
        np.savez_compressed(os.path.join(G.GOLD, "c2.npz"), **out)
    elif what == "c2planes":
        # At 1.6e8 values per array a float64 rounding difference (LAPACK gbsv vs a streaming Thomas recurrence, FMA vs
        # mul + add) is expected to flip the float32 rounding of a handful of values, so a whole-array SHA-256 is too
        # brittle at this size: per-x-plane CRC32s of the flushed arrays localise and count the differing planes.
        from mad.MapSpace import MapSpace
        grid, _ = synth.c2_inputs(0)
        mrc = os.path.join(G.WORK, "c2map.mrc")
        ref_shims.write_mrc_stub(mrc, np.ascontiguousarray(grid, dtype=np.float32), synth.C2["voxelsp"], (0.0, 0.0, 0.0))
        ms = MapSpace(mrc)
        ms.build_space()
        path = os.path.join(G.GOLD, "c2.npz")
        with np.load(path, allow_pickle=False) as z:
            out = {k: z[k] for k in z.files}
        arrays = {"up_grid": ms.grid_list[0]}
        for o in range(2):
            arrays["log%d" % o], arrays["gauss%d" % o], arrays["grad%d" % o] = ms.map_space[o], ms.gauss_list[o], ms.grad_list[o]
        for k, a in arrays.items():
            assert G.sha_flushed(a) == str(out[k + "_sha256_flushed"]), k
            out[k + "_plane_crc32_flushed"] = plane_crcs(a)
        np.savez_compressed(path, **out)
        np.save(os.path.join(PARTS, "c2_up_grid.npy"), ms.grid_list[0])
        print("added plane CRCs to c2.npz")
    elif what == "c2comp":
        _, comps = synth.c2_inputs(0)
        for i in [int(a) for a in argv[1:]]:
            out, dsc = run_map("c2comp%d" % i, comps[i], synth.C2["voxelsp"])
            np.save(os.path.join(PARTS, "c2comp%d_dsc.npy" % i), dsc)
            np.savez_compressed(os.path.join(PARTS, "c2comp%d.npz" % i), **out)
    elif what == "c2pairs":
        merged = {}
        his = []
        for i in range(synth.C2["n_sub"]):
            with np.load(os.path.join(PARTS, "c2comp%d.npz" % i)) as z:
                for k in z.files:
                    if k.startswith(("kp_", "of_", "dsc_", "input_", "ref_timings")) or k.endswith("_sha256_flushed"):
                        merged["comp%d_%s" % (i, k)] = z[k]
            his.append(np.load(os.path.join(PARTS, "c2comp%d_dsc.npy" % i)))
        lo = np.load(os.path.join(PARTS, "c2map_dsc.npy"))
        hi = np.concatenate(his)
        cc = 0.6
        hu, lu = unit_rows(hi), unit_rows(lo)
        crc, count, ssum, margin = 0, 0, 0.0, np.inf
        per_row = np.zeros(len(hi), dtype=np.int32)
        row_crc = np.zeros(len(hi), dtype=np.uint32)
        for s in range(0, len(hi), 2048):
            preds = np.dot(hu[s:s + 2048], lu.T)                                  # mad/MaD.py:420
            pairs = np.array(np.where(preds > cc)).T.astype(np.int32)             # mad/MaD.py:423-424
            sc = preds[pairs[:, 0], pairs[:, 1]]
            margin = min(margin, float(np.abs(preds - cc).min()))
            pairs[:, 0] += s
            crc = zlib.crc32(np.ascontiguousarray(pairs).tobytes(), crc)
            count += len(pairs)
            ssum += float(sc.sum())
            np.add.at(per_row, pairs[:, 0], 1)
            starts = np.searchsorted(pairs[:, 0], np.arange(s, min(s + 2048, len(hi)) + 1))
            for r in range(len(starts) - 1):
                row_crc[s + r] = zlib.crc32(np.ascontiguousarray(pairs[starts[r]:starts[r + 1], 1]).tobytes())
            print("pairs: rows %d / %d, %d so far" % (s, len(hi), count), flush=True)
        merged.update(pairs_cc=np.array(cc), pairs_count=np.array(count, dtype=np.int64), pairs_crc32=np.array(crc, dtype=np.uint32),
                      pairs_per_hi_row=per_row, pairs_lo_crc32_per_hi_row=row_crc, pairs_score_sum=np.array(ssum),
                      pairs_min_margin=np.array(margin), hi_rows=np.array(len(hi)), lo_rows=np.array(len(lo)),
                      comp_rows=np.array([len(h) for h in his], dtype=np.int64))
        np.savez_compressed(os.path.join(G.GOLD, "c2_comp.npz"), **merged)
        print("wrote c2_comp.npz: %d x %d, %d pairs, closest score to cc %.3g" % (len(hi), len(lo), count, margin))
    elif what == "c4":
        ids = [int(a) for a in argv[1:]]
        base = synth.random_walk_atoms(9000, 85.0, 1)
        merged = {"snapshots": np.array(ids, dtype=np.int64)}
        for i in ids:
            out, _ = run_map("c4snap%d" % i, synth.c4_snapshot(i, base), 1.0)
            for k, v in out.items():
                if k.startswith(("kp_", "of_", "dsc_", "input_", "ref_timings")) or k.endswith("_sha256_flushed"):
                    merged["s%d_%s" % (i, k)] = v
        np.savez_compressed(os.path.join(G.GOLD, "c4.npz"), **merged)
    elif what == "c5":
        m = n = 100000
        hi, lo = synth.c5_descriptor_sets(m, n)
        rows = np.sort(np.random.default_rng(5).choice(m, size=256, replace=False))
        preds = np.dot(unit_rows(hi[rows]), unit_rows(lo).T)                      # mad/MaD.py:420 on the sampled rows
        order = np.argsort(-preds, axis=1, kind="stable")[:, :9]                  # stable top-k (9th = tie witness)
        sc = np.take_along_axis(preds, order, 1)
        np.savez_compressed(os.path.join(G.GOLD, "c5.npz"), rows=rows.astype(np.int64), top9_idx=order.astype(np.int32),
                            top9_score=sc, m=np.array(m), n=np.array(n),
                            hi_sha256=np.array(G.sha(hi)), lo_sha256=np.array(G.sha(lo)),
                            pairs_over_cc=np.array([(preds > 0.6).sum(1)], dtype=np.int64).reshape(-1))
        print("wrote c5.npz")
    else:
        raise SystemExit("unknown job %r" % what)


if __name__ == "__main__":
    main(sys.argv[1:])
