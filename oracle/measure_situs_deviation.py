"""TEST INFRASTRUCTURE ONLY -- measures what the float32 treatment of Situs (.sit) inputs changes.

The reference keeps a .sit grid in float64 (np.fromstring -> float64, `grid / np.float32(max)` stays float64,
mad/MapSpace.py:90-96), so its base-octave filters run on float64 input; this package's reader casts to float32 like the
MRC branch.  Here the UNMODIFIED reference runs the `small` case twice -- from a .sit file (float64 path) and from an MRC
file holding the same parsed values as float32 -- and the sparse results are key-joined.  Needs /root/reference.

    python oracle/measure_situs_deviation.py  > profiles/r02_situs_float64_deviation.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_goldens as G  # noqa: E402
import ref_shims  # noqa: E402
import synth  # noqa: E402


def main():
    G._enter_workdir()
    with np.load(os.path.join(G.GOLD, "small.npz"), allow_pickle=False) as z:
        grid = synth.dequantise_u16(z["input_q"])
        v = float(z["voxelsp"])
    sit = os.path.join(G.WORK, "dev.sit")
    xb, yb, zb = grid.shape
    with open(sit, "w") as f:
        f.write("%f %f %f %f %i %i %i\n\n" % (v, 0.0, 0.0, 0.0, xb, yb, zb))
        k = 0
        for zz in range(zb):
            for yy in range(yb):
                for xx in range(xb):
                    f.write("   %6.6f " % grid[xx][yy][zz])
                    k += 1
                    if k % 10 == 0:
                        f.write("\n")
    ms_s, a_s, d_s, _ = G.run_pipeline(sit)
    # the same values, as the float32 grid this package's Situs reader produces, through the MRC branch
    with open(sit) as f:
        f.readline(); f.readline()
        vals = np.array(f.read().split(), dtype=np.float64)
    g64 = np.reshape(vals, (xb, yb, zb), order="F")
    g32 = (g64 / np.amax(g64).astype(np.float32)).astype(np.float32)
    mrc = os.path.join(G.WORK, "dev.mrc")
    ref_shims.write_mrc_stub(mrc, g32, v, (0.0, 0.0, 0.0))
    ms_m, a_m, d_m, _ = G.run_pipeline(mrc)

    def keys(anchors):
        return {(a.oct_scale,) + tuple(int(c) for c in a.coords) for a in anchors}

    def dkeys(desc):
        return {((d.oct_scale,) + tuple(int(c) for c in d.coords) + (d.main_bin, d.sec_bin)): np.asarray(d.lin_ar_subeqsp) for d in desc}
    ks, km = keys(a_s), keys(a_m)
    ds, dm = dkeys(d_s), dkeys(d_m)
    common = set(ds) & set(dm)
    out = dict(case="small", situs_dtype=str(ms_s.grid_list[1].dtype), mrc_dtype=str(ms_m.grid_list[1].dtype),
               keypoints_situs=len(ks), keypoints_float32=len(km), keypoint_flips=len(ks ^ km),
               oriented_situs=len(ds), oriented_float32=len(dm), oriented_flips=len(set(ds) ^ set(dm)),
               descriptor_rows_differing=int(sum(1 for c in common if not np.array_equal(ds[c], dm[c]))),
               log_base_max_abs_diff=float(np.abs(ms_s.map_space[1].astype(np.float64) - ms_m.map_space[1]).max()))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
