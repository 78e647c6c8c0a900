"""Kernel time of the C2 matching launch as a function of the threshold (= hit density).  python scripts/time_match_c2.py"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import synth  # noqa: E402
import bench  # noqa: E402
from mad_b200 import pipeline as P  # noqa: E402

grid_h, comps_h = synth.assembly_with_components(**dict(bench.C2))
sets = []
for c in comps_h:
    _, _, _, dsc = P.describe_struct(c)
    sets.append(P.DescriptorSet(dsc))
hi_all, _ = P.concat_sets(sets)
_, _, _, dsc = P.describe_struct(torch.from_numpy(grid_h).cuda())
lo = P.DescriptorSet(dsc)
for cc in (0.99, 0.8, 0.7, 0.65, 0.6, 0.55):
    for _ in range(2):
        ph, pl, sc = P.match_threshold(hi_all, lo, cc)
    torch.cuda.synchronize()
    P.profile_enable(True)
    for _ in range(5):
        ph, pl, sc = P.match_threshold(hi_all, lo, cc)
    torch.cuda.synchronize()
    recs = P.profile_records()
    P.profile_enable(False)
    tot = {}
    for nm, t in recs:
        tot[nm] = tot.get(nm, 0.0) + t / 5
    print("cc=%.2f pairs=%8d  " % (cc, ph.numel()) + "  ".join("%s=%.3f" % (k.replace("_kernel", ""), v) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:3]))
