"""The matching launch of the C2 bench step (six component descriptor sets stacked against the map's) inside a
cudaProfilerStart/Stop range, for ncu --profile-from-start off.  Usage: python scripts/profile_match_c2.py"""
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
for _ in range(2):
    ph, pl, sc = P.match_threshold(hi_all, lo, 0.6)
torch.cuda.synchronize()
print("M=%d N=%d pairs=%d" % (hi_all.rows, lo.rows, ph.numel()))
torch.cuda.profiler.start()
P.match_threshold(hi_all, lo, 0.6)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
